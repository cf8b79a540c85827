// conv_tc.cu -- tcgen05 implicit-GEMM convolution for the discriminator forward (the throughput path).
//
// Replaces F.conv2d(x, W/sigma, b, stride 1, pad ks/2) of torch-mimicry SNConv2d as called by the
// SNGAN discriminator blocks (call site trainer.py:150) for 3x3 and 1x1 kernels.
//
// GEMM view: D[M = pixels, N = Cout] = A[M, K = taps*Cin] * B[N, K]^T, bf16 operands, fp32 accumulate.
//   * A is never materialised: NHWC activations are read by 4-D TMA boxes {64 channels, bw, bh, bn}
//     (bw*bh*bn = 128 consecutive pixels of full-width rows), one box per (tap, 64-channel chunk),
//     shifted by the tap offset; the halo of the 3x3 window is the TMA out-of-bounds zero fill.
//     A box lands in shared memory as 128 rows x 128 B, 128B-swizzled = the canonical K-major UMMA
//     operand layout.
//   * B (weights, [Cout][taps*Cin] bf16, sigma folded in) is read by 2-D TMA boxes {64, BN}.
//   * one elected thread issues tcgen05.mma (M=128, N=BN, K=16) into a double-buffered TMEM
//     accumulator; stages are recycled with tcgen05.commit -> mbarrier.
//   * epilogue warps read TMEM with tcgen05.ld, add bias, optional ReLU, pack bf16, store NHWC.
// Persistent CTAs (one per SM), static tile schedule, warp-specialised: warp 0 TMA producer, warp 1
// MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue.
#include <cuda.h>

#include "kernels.cuh"

namespace sdg {

constexpr int TC_BM = 128;            // pixels per tile (UMMA M)
constexpr int TC_BK = 64;             // channels per stage (128 B of bf16 = one swizzle row)
constexpr int TC_STAGES = 6;
constexpr int TC_THREADS = 256;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;     // 16 KB
constexpr int TC_TMEM_COLS = 256;                 // 2 accumulator stages x up to 128 fp32 columns
constexpr int TC_MAX_COUT = 1024;

struct TcParams {
  int H, W, Cin, Cout, taps;
  int bh, tiles_y, bn;       // tile = bn images x bh rows x W columns
  int n_tiles;               // Cout / BN
  int kchunks;               // Cin / 64
  int post_relu;
  long long m_tiles;
  long long total_pixels;
};

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug must fail the launch (trap), never hang the GPU box
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 6000000000LL) {      // ~3 s at 2 GHz
      printf("sdg conv_tc: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc]^T, kind::f16 (bf16 inputs, fp32 accumulate)
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart (SBO); LBO unused
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address, 16-byte units        bits [0,14)
  d |= (uint64_t)1 << 16;                             // leading byte offset (ignored)       bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset                  bits [32,46)
  d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)      bits [46,48)
  d |= (uint64_t)2 << 61;                             // layout type SWIZZLE_128B            bits [61,64)
  return d;
}

// instruction descriptor: D fp32 (bits 4-5 = 1), A/B format (bits 7-9 / 10-12: 0 = fp16, 1 = bf16),
// both operands K-major, N >> 3 at bits 17-22, M >> 4 at bits 24-28
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool f16) {
  return (1u << 4) | ((f16 ? 0u : 1u) << 7) | ((f16 ? 0u : 1u) << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------------
template <int BN, bool F16>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const float* __restrict__ bias, h16* __restrict__ out, const TcParams p) {
  constexpr int B_BYTES = BN * TC_BK * 2;
  constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment: required by the 128B swizzle atom (TMA and UMMA must agree on address bits)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  __shared__ __align__(8) uint64_t bar_full[TC_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[TC_STAGES];
  __shared__ __align__(8) uint64_t bar_acc_full[2];
  __shared__ __align__(8) uint64_t bar_acc_empty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_bias[TC_MAX_COUT];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < p.Cout; i += TC_THREADS) s_bias[i] = bias ? bias[i] : 0.f;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
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
  if (warp == 2) tmem_alloc(smem_u32(&tmem_base_slot), TC_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const long long total_tiles = p.m_tiles * p.n_tiles;
  const int k_iters = p.taps * p.kchunks;

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = (int)(tile % p.n_tiles);
        const long long mt = tile / p.n_tiles;
        const int ty = (int)(mt % p.tiles_y);
        const int n0 = (int)(mt / p.tiles_y) * p.bn;
        const int y0 = ty * p.bh;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dy = p.taps == 9 ? tap / 3 - 1 : 0;
          const int dx = p.taps == 9 ? tap % 3 - 1 : 0;
          for (int kc = 0; kc < p.kchunks; ++kc) {
            mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
            const uint32_t full = smem_u32(&bar_full[stage]);
            mbar_expect_tx(full, STAGE_BYTES);
            const uint32_t a_dst = smem_base + stage * STAGE_BYTES;
            tma_load_4d(a_dst, &map_a, full, kc * TC_BK, dx, y0 + dy, n0);
            tma_load_2d(a_dst + TC_A_BYTES, &map_b, full, tap * p.Cin + kc * TC_BK, nt * BN);
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
          }
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
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * STAGE_BYTES;
          const uint64_t adesc = make_sw128_desc(a_addr);
          const uint64_t bdesc = make_sw128_desc(a_addr + TC_A_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            // advance 16 bf16 = 32 B inside the swizzle row: +2 in the 16-byte-unit address field
            umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (it | k) != 0 ? 1u : 0u);
          }
          umma_commit(smem_u32(&bar_empty[stage]));       // frees the smem stage when these MMAs retire
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(smem_u32(&bar_acc_full[acc]));        // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue: TMEM -> registers -> bias/ReLU -> bf16 -> global (NHWC) =================
    const int q = warp - 4;                       // TMEM lane quadrant this warp may access
    long long local = 0;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int nt = (int)(tile % p.n_tiles);
      const long long mt = tile / p.n_tiles;
      const int acc = (int)(local & 1);
      const uint32_t acc_phase = (uint32_t)((local >> 1) & 1);
      mbar_wait(smem_u32(&bar_acc_full[acc]), acc_phase);
      tc_fence_after();
      const long long pix = mt * TC_BM + q * 32 + lane;      // tile rows are 128 consecutive NHW pixels
      const bool valid = pix < p.total_pixels;
      h16* orow = out + pix * p.Cout + nt * BN;
      const float* brow = s_bias + nt * BN;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), r);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 pk;
            uint32_t* h = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float x = __uint_as_float(r[g * 8 + 2 * j]) + brow[c0 + g * 8 + 2 * j];
              float y = __uint_as_float(r[g * 8 + 2 * j + 1]) + brow[c0 + g * 8 + 2 * j + 1];
              if (p.post_relu) { x = fmaxf(x, 0.f); y = fmaxf(y, 0.f); }
              h[j] = pack_h2<F16>(x, y);
            }
            *reinterpret_cast<uint4*>(orow + c0 + g * 8) = pk;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[acc]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TC_TMEM_COLS);
  }
  (void)smem_gen;
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

int conv_tc_init(int device) {
  if (g_encode) return 0;
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
  g_encode = (EncodeTiledFn)fn;
  return 0;
}

int conv_tc(const h16* in, const h16* wb, const float* bias, h16* out, int64_t n, int H, int W, int Cin, int Cout,
            int taps, int post_relu, int f16, cudaStream_t s) {
  SDG_REQUIRE(g_encode, SDG_E_STATE, "conv_tc: conv_tc_init not called");
  SDG_REQUIRE(taps == 9 || taps == 1, SDG_E_UNSUPPORTED, "conv_tc: taps=%d", taps);
  SDG_REQUIRE(Cin % TC_BK == 0 && Cout % 64 == 0 && Cout <= TC_MAX_COUT, SDG_E_UNSUPPORTED, "conv_tc: Cin=%d Cout=%d", Cin, Cout);
  SDG_REQUIRE(W >= 4 && W <= 128 && (W & (W - 1)) == 0 && H == W, SDG_E_UNSUPPORTED, "conv_tc: H=%d W=%d", H, W);
  SDG_REQUIRE(((uintptr_t)in % 16) == 0 && ((uintptr_t)wb % 16) == 0 && ((uintptr_t)out % 16) == 0, SDG_E_INVALID,
              "conv_tc: pointers must be 16-byte aligned");
  if (n == 0) return 0;
  TcParams p;
  p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.taps = taps; p.post_relu = post_relu;
  int rows = TC_BM / W;                       // image rows per tile if one image is big enough
  if (rows >= H) { p.bh = H; p.bn = TC_BM / (H * W); } else { p.bh = rows; p.bn = 1; }
  p.tiles_y = H / p.bh;
  const int BN = (Cout % 128 == 0) ? 128 : 64;
  p.n_tiles = Cout / BN;
  p.kchunks = Cin / TC_BK;
  p.m_tiles = cdiv(n, p.bn) * p.tiles_y;
  p.total_pixels = n * H * W;

  CUtensorMap map_a, map_b;
  const CUtensorMapDataType dtype = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
    cuuint32_t box[4] = {(cuuint32_t)TC_BK, (cuuint32_t)W, (cuuint32_t)p.bh, (cuuint32_t)p.bn};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode(&map_a, dtype, 4, (void*)in, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SDG_REQUIRE(r == CUDA_SUCCESS, SDG_E_INVALID, "conv_tc: cuTensorMapEncodeTiled(A) failed: %d", (int)r);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)taps * Cin, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)taps * Cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(&map_b, dtype, 2, (void*)wb, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SDG_REQUIRE(r == CUDA_SUCCESS, SDG_E_INVALID, "conv_tc: cuTensorMapEncodeTiled(B) failed: %d", (int)r);
  }
  const long long total_tiles = p.m_tiles * p.n_tiles;
  const int grid = (int)(total_tiles < g_num_sms ? total_tiles : g_num_sms);
  if (BN == 128 && f16) {
    SDG_LAUNCH((conv_tc_kernel<128, true>), grid, TC_THREADS, tc_smem_bytes<128>(), s, map_a, map_b, bias, out, p);
  } else if (BN == 128) {
    SDG_LAUNCH((conv_tc_kernel<128, false>), grid, TC_THREADS, tc_smem_bytes<128>(), s, map_a, map_b, bias, out, p);
  } else if (f16) {
    SDG_LAUNCH((conv_tc_kernel<64, true>), grid, TC_THREADS, tc_smem_bytes<64>(), s, map_a, map_b, bias, out, p);
  } else {
    SDG_LAUNCH((conv_tc_kernel<64, false>), grid, TC_THREADS, tc_smem_bytes<64>(), s, map_a, map_b, bias, out, p);
  }
  return 0;
}

}  // namespace sdg
