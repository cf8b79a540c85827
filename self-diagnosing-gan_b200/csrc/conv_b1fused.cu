// conv_b1fused.cu -- SNGAN-32 block 1 (DBlockOptimized) as ONE kernel: the 256 KB / sample tensor relu(c1(x)) never leaves the SM.
//
// Replaces, for torch-mimicry's DBlockOptimized(3, 128) of SNGANDiscriminator32 (call site trainer.py:150):
//     T   = relu(conv3x3(normalise(x)) + b1)                        (first_conv_kernel: 3 KB in, 256 KB out per sample)
//     out = relu(avg_pool2d(conv3x3(T) + b2) + Wsc . avg_pool2d(x) + bsc)   (conv_swap_shared_kernel, 4x4 stride-2 form:
//                                                                            256 KB in, 64 KB out per sample)
// Unfused, T is the largest stream of the whole pass (HBM write-bound first conv: 17 % of the SNGAN-32 forward) and its re-read
// is what makes block1.c2 operand-ingest bound.  Here a CTA PAIR owns one image (CTA rank s = the left / right strip of 8 output
// columns), builds its strip of T in shared memory with the first conv's own arithmetic (same LUT-normalised 16-bit operand,
// same K = 32 tcgen05.mma with the bias in columns 27 / 28, same cvt.rn.relu rounding: T is bit-identical to first_conv's),
// and runs the 4x4 stride-2 form of c2 as tcgen05.mma.cta_group::2 with M = 256 pixels (128 per CTA) x N = 128 channels
// straight from that tile.  Only the weights stream (8 KB per CTA per 256 MMA cycles = 32 B per cycle).
//
// T in shared memory uses the NO-SWIZZLE K-major UMMA layout: 16-byte cells (8 channels of one pixel); 8 consecutive M rows
// 16 bytes apart, groups of 8 rows SBO apart, the two K halves of an instruction LBO apart.  With SBO = one plane row the 8-row
// groups are the tile's 16 output rows, so a tap's ROW shift and its COLUMN shift are both just a different descriptor start
// address (the 128B-swizzled layout cannot express a one-pixel column shift; that is why the unfused kernel needs one staged
// copy per column variant, and why a resident T did not fit before: profiles/r2b_swap_kernel_limiters.md section 4).
//   T cell (ty, tx, k8):  ty = 2R + pr in 0..33 (image row + 1), tx = 2C + pc in 0..17 (image column - 16 s + 1)
//       byte offset ((pr*2 + pc)*16 + k8) * 2448 + R * 144 + C * 16            4 parity planes x 16 x (17 rows x 9 cells) = 153 KB
//   tap (ky, kx) of output pixel (r, c): plane (ky & 1, kx & 1), R = r + (ky >> 1), C = c + (kx >> 1).
// The halo cells (ty = 0, 33; tx = 0 for s = 0, tx = 17 for s = 1) are zeroed once and never written.
//
// Pipeline (per CTA; all barriers as in conv_pair_stream_kernel: "full"-type barriers live in the leader and are signalled by
// both CTAs, "empty"-type ones are multicast by the leader's tcgen05.commit):
//   unit = one ROW-PARITY half of T (272 pixels): c2's taps ky in {1, 3} read only odd ty, ky in {0, 2} only even ty, so while
//   the tensor pipe runs the 64 MMAs of one half, the T warps rebuild the other half for the next image: no second T buffer.
//   T warps (4..7 and 12..15): gather 128 pixels x 27 bytes -> A1 (no-swizzle im2col tile, double-buffered) -> [c1 MMA, K = 32]
//   -> TMEM -> relu/convert -> T cells; three batches per unit (128 + 128 + 16 pixels); the two sets split the K columns of the
//   gather and the channels of the drain.  Epilogue warps (8..11): TMEM -> + bias + 3-FMA image shortcut -> relu -> 16-bit NHWC.
//   Warp 0: weight TMA, warp 1: c2 MMA issuer, warp 3: c1 MMA issuer (leader CTA), warp 2: TMEM.
//
// CH = 64 (SNGANDiscriminator64's DBlockOptimized(3, 64), 64 x 64 images): the same machine per QUADRANT of an image.  A tile
// is the 32 x 32 pixel quadrant (qy, qx) = 16 x 16 pooled pixels; its T strip (34 x 18 cells per CTA) now carries real halo
// values -- relu(c1(x)) of the neighbouring quadrants' border pixels, recomputed here (c1 is 2 % of the MACs), zeros only
// outside the image -- so every cell of a unit is rebuilt for every tile and the patch spans image rows / columns -2 .. 33 of
// the quadrant.  N = 64: tcgen05.mma.cta_group::2 of M = 256 pixels x N = 64 channels, no structural zeros (the unfused
// path packs two pooled pixels into one 128-channel GEMM pixel and spends a third of its MACs on zeros); a 64-channel tap is
// one K chunk, and a weight stage holds TWO taps (the same 8 KB per CTA and 256 MMA cycles per barrier as CH = 128).
// A 64-channel c2 needs a tile's T four times sooner than the 128-channel one, so here the T warps -- not the tensor pipe --
// pace the kernel, and behind them the shared-memory data pipe (N = 64 MMAs fetch 5 KB of operands per 32 tensor cycles).
// Hence for CH = 64: (1) the T warps only gather and drain; the landing rows, the normalised patch and the shortcut operand of
// a tile are produced by the EPILOGUE warps, two tiles ahead, into double buffers (mbarrier hand-off, no block-wide barrier);
// (2) the two T-warp sets take alternate batches with two im2col buffers / accumulators each; (3) one small copy of the T loop
// with all data-independent arithmetic hoisted (instruction-cache footprint matters: a 32 KB specialised loop ran 40 % slower);
// (4) bank-conflict-free shared-memory traffic: the patch as two column-parity planes with 96-byte rows, lanes mapped to
// 8-cell rows, word loads + arithmetic instead of byte loads + table lookups.  profiles/r4_b1fused64.md has the measurements.
#include <cstdlib>
#include <type_traits>

#include "tc_ptx.cuh"

namespace sdg {

constexpr int BF_THREADS = 512;
constexpr int BF_T_THREADS = 256;                       // two sets of four T warps (4..7 and 12..15)
constexpr int BF_W_MAX = 8;                             // barrier slots = deepest weight ring
constexpr int BF_W_BYTES = 64 * 64 * 2;                 // 8 KB per stage: CH = 128: this CTA's 64 output channels x 64 k (half a tap);
                                                        // CH = 64: its 32 channels x 64 k of two taps (2 x 4 KB)
constexpr int BF_ROW = 9 * 16;                          // 144: one plane row = 9 cells
constexpr int BF_K8 = 17 * BF_ROW;                      // 2448: 17 plane rows per 8-channel group
constexpr int BF_A1_BYTES = 4 * 128 * 16;               // 8192: [k8 0..3][128 pixels] x 16 B
constexpr int BF_X_ROWB = 64;                           // raw image rows: the bytes of the strip's columns (CH = 128: 56 at offset 4)
constexpr int BF_SC_BYTES = 2 * 128 * 16;               // shortcut operand: [k8 0..1][128 output pixels] x 16 B
// CH = 64: the patch is split into its even and odd COLUMNS (two planes of 36 rows x 10 pixels): the three taps of a row are then
// at (plane pc, C), (plane 1 - pc, C + pc), (plane pc, C + 1) with consecutive lanes (C) 8 bytes apart, and with 96-byte rows
// two rows of 8 pixels fill the 32 banks exactly
constexpr int BF_PQ_ROWB = 96;
constexpr int BF_PQ_PLANE = 36 * BF_PQ_ROWB;
constexpr int BF_P_ROWB = 160;                          // CH = 128: one interleaved plane, 34 rows x 20 pixels x (c0, c1, c2, pad): the column-parity
                                                        // split was measured there too -- 4.8 % MORE cycles per image under ncu (store replays
                                                        // +200 per image) and 2 % on a 6 250-sample pass, no change on the power-capped 50 k pass

template <int CH>
struct Bf {
  static constexpr bool QUAD = CH == 64;                // tile = quadrant of a 64 x 64 image (else: the 32 x 32 image)
  static constexpr int IMG = QUAD ? 64 : 32;            // image side
  static constexpr int NH = CH / 2;                     // output channels (B rows) per CTA
  static constexpr int G = CH / 8;                      // 8-channel groups per pixel
  static constexpr int PLANE = G * BF_K8;               // bytes per (row parity, column parity) plane: 39168 | 19584
  static constexpr int T_BYTES = 4 * PLANE;             // 156672 | 78336
  static constexpr int W1_BYTES = 4 * NH * 16;          // [k8 0..3][NH channels] x 16 B
  static constexpr int W_STAGES = QUAD ? 8 : 5;         // weight ring depth (T leaves room for 8 when CH = 64)
  static constexpr int CHUNKS = QUAD ? 4 : 16;          // weight stages per phase (8 taps: half a tap | two taps per stage)
  static constexpr int X_ROWS = QUAD ? 36 : 32;         // landing buffer: image rows -2 .. 33 of the quadrant | the image's 32 rows
  static constexpr int X_BYTES = X_ROWS * BF_X_ROWB;
  static constexpr int P_ROWS = QUAD ? 36 : 34;         // patch rows -2 .. 33 | -1 .. 32
  static constexpr int P_BYTES = QUAD ? 2 * BF_PQ_PLANE : P_ROWS * BF_P_ROWB;
  static constexpr int OFF_T = W_STAGES * BF_W_BYTES;
  static constexpr int OFF_A1 = OFF_T + T_BYTES;
  static constexpr bool ALT = QUAD;                     // the two T-warp sets take alternate batches (below) instead of halves of each
  static constexpr int NA1 = ALT ? 6 : 2;               // im2col buffers = c1 accumulators: three per set | double-buffered
  static constexpr int OFF_W1 = OFF_A1 + NA1 * BF_A1_BYTES;
  static constexpr bool EPI_PATCH = QUAD;               // the patch / shortcut operand of a tile is built by the EPILOGUE warps, two
                                                        // tiles ahead, instead of by the T warps (which pace the CH = 64 kernel)
  static constexpr int NB = EPI_PATCH ? 2 : 1;          // ... into double-buffered landing rows, patch and shortcut operand
  static constexpr int OFF_X = OFF_W1 + W1_BYTES;
  static constexpr int OFF_P = OFF_X + NB * X_BYTES;
  static constexpr int OFF_SC = OFF_P + NB * P_BYTES;
  static constexpr int OFF_END = OFF_SC + NB * BF_SC_BYTES;
  static constexpr int SMEM = 1024 + OFF_END;
  static constexpr int W2_LD = 16 * CH + 64;            // packed c2 weights: 16 taps x CH channels + one 64-column chunk for the shortcut
  static constexpr int SC_COL = 16 * CH;                // ... which starts at this column
  static constexpr int UNIT_CELLS = QUAD ? 306 : 272;   // T cells one unit (row-parity half) computes: 2 x 17 x 9 | 16 x 8 + 16 x 9
  static_assert(SMEM + 512 <= 227 * 1024, "b1_fused_kernel: shared memory budget (static: < 512 B of barriers)");
};

struct BfParams {
  const uint8_t* x;          // [n][IMG][IMG][3]
  const h16* w1;             // [CH][64] K-major, k = (ky*3+kx)*3 + c (27 real columns)
  const float* b1;           // [CH]
  h16* out_relu;             // [n][IMG/2][IMG/2][CH]
  h16* dbg_t;                // [n][IMG][IMG][CH] copy of T (tests only) or null
  int* ovf;                  // fp16 range guard flag or null
  long long n_tiles;         // images (CH = 128) or image quadrants (CH = 64: tile = 4 * image + 2 * qy + qx)
  int w_stages;              // weight ring depth in use (2..Bf<CH>::W_STAGES).  tcgen05.mma executes in issue order, so the ring depth also
                             // bounds how many c2 chunks (256 MMA cycles each) can be queued ahead of a c1 batch.
                             // (SDG_TIMING_EXPERIMENTS builds accept deeper rings that overwrite T: wrong results)
  int dbg;                   // SDG_TIMING_EXPERIMENTS builds only (WRONG results): 1 no gather, 2 no T drain, 4 no c2 epilogue, 8 no c1 MMAs,
                             // 16 no c2 MMAs, 32 no weight loads
};

// normalise as transform.py:3-11 does in fp32, (u / 255 - 0.5) / 0.5, without the two IEEE divisions: u * (1/255) with one
// Newton correction is the correctly rounded quotient for every u in 0..255 (checked exhaustively; the bit-identity test of
// relu(c1(x)) against first_conv_kernel, which divides, covers it on the GPU), and dividing by 0.5 is an exact doubling
__device__ __forceinline__ float bf_nrm(uint8_t u) {
  const float f = (float)u, r = 1.0f / 255.0f;
  float q = __fmul_rn(f, r);
  q = __fmaf_rn(__fmaf_rn(-q, 255.0f, f), r, q);
  return __fmul_rn(__fsub_rn(q, 0.5f), 2.0f);
}

// K-major operand without swizzle: rows of a group 16 B apart, 8-row groups `sbo` bytes apart, K halves `lbo` bytes apart
__device__ __forceinline__ uint64_t make_nosw_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;                                             // layout type 0 = SWIZZLE_NONE
}

// weight chunk `idx` (0..15) of phase `ph`: ph 0 = taps ky in {1, 3} (odd T rows), ph 1 = ky in {0, 2}; kx 0..3; h = 64-channel half
__device__ __forceinline__ void bf_chunk(int ph, int idx, int& ky, int& kx, int& h) {
  ky = (ph == 0 ? 1 : 0) + 2 * (idx >> 3);
  kx = (idx >> 1) & 3;
  h = idx & 1;
}

// pixel `m` of batch `b` of the unit with row parity `pr`, strip `s`: cell (R, C) of column-parity plane pc
template <bool QUAD>
__device__ __forceinline__ bool bf_pixel(int s, int pr, int b, int m, int& R, int& C, int& pc) {
  if (QUAD) {
    // every cell of both column-parity planes, 2 x (17 rows x 9 cells): first cells 0..7 of every row (a warp = 4 rows x 8 cells:
    // its LDS.64 gathers from the patch and its STS.128 to T are bank-conflict free), then the ninth cell of the 2 x 17 rows
    const int idx = b * 128 + m;
    if (idx < 272) {
      pc = idx >= 136 ? 1 : 0;
      const int rem = idx - 136 * pc;
      R = rem >> 3; C = rem & 7;
    } else {
      pc = idx >= 289 ? 1 : 0;
      R = idx - 272 - 17 * pc; C = 8;
    }
    return idx < 306;
  }
  int rr;
  bool valid = true;
  if (b == 0) {                                         // the 8-column parity plane: 16 rows x 8 cells
    rr = m >> 3; C = (m & 7) + (1 - s); pc = s;
  } else {                                              // the 9-column one: 144 cells in address order
    const int idx = (b - 1) * 128 + m;
    valid = idx < 144;
    rr = idx / 9; C = idx - rr * 9; pc = 1 - s;
  }
  R = rr + (pr ? 0 : 1);
  return valid;
}

template <bool F16, int CH>
__global__ void __launch_bounds__(BF_THREADS, 1)
b1_fused_kernel(const __grid_constant__ CUtensorMap map_w2, const BfParams p) {
  using B = Bf<CH>;
  constexpr bool QUAD = B::QUAD, ALT = B::ALT, EPI_PATCH = B::EPI_PATCH;
  static_assert(!EPI_PATCH || (QUAD && ALT), "the epilogue-built patch is written for the CH = 64 schedule");
  constexpr int NH = B::NH, G = B::G, IMG = B::IMG;
  constexpr int BF_OFF_T = B::OFF_T, BF_OFF_A1 = B::OFF_A1, BF_OFF_W1 = B::OFF_W1, BF_OFF_X = B::OFF_X, BF_OFF_P = B::OFF_P,
                BF_OFF_SC = B::OFF_SC, BF_T_BYTES = B::T_BYTES, BF_P_BYTES = B::P_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  __shared__ __align__(8) uint64_t bar_wfull[BF_W_MAX];
  __shared__ __align__(8) uint64_t bar_wempty[BF_W_MAX];
  __shared__ __align__(8) uint64_t bar_a1_full[6];
  __shared__ __align__(8) uint64_t bar_c1_full[6];
  __shared__ __align__(8) uint64_t bar_c1_empty[6];
  __shared__ __align__(8) uint64_t bar_t_ready[2];
  __shared__ __align__(8) uint64_t bar_t_free[2];
  __shared__ __align__(8) uint64_t bar_acc_full[2];
  __shared__ __align__(8) uint64_t bar_acc_empty[2];
  __shared__ __align__(8) uint64_t bar_sc_full[2];      // shortcut operand of a tile built (both CTAs) -> issuer   ([1]: EPI_PATCH only)
  __shared__ __align__(8) uint64_t bar_sc_free[2];      // its MMA has retired -> the buffer may be overwritten
  __shared__ __align__(8) uint64_t bar_patch_ready[2];  // EPI_PATCH: patch buffer (tile & 1) written (this CTA's epilogue warps) -> T warps
  __shared__ __align__(8) uint64_t bar_patch_free[2];   // ... and every T warp of this CTA has gathered its last batch of that tile from it
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int s = (int)rank;                              // strip: output columns 8s .. 8s+7
  const bool leader = rank == 0;
#ifdef SDG_TIMING_EXPERIMENTS
  const int dbg = p.dbg;
#else
  constexpr int dbg = 0;
#endif
  const int w_stages = p.w_stages;

  // ---- one-time setup ----
  for (int i = threadIdx.x; i < BF_T_BYTES / 16; i += BF_THREADS)
    reinterpret_cast<uint4*>(smem_gen + BF_OFF_T)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < B::NB * BF_P_BYTES / 16; i += BF_THREADS)
    reinterpret_cast<uint4*>(smem_gen + BF_OFF_P)[i] = make_uint4(0u, 0u, 0u, 0u);
  // c1 weights of this CTA's 64 channels, no-swizzle [k8][row]; the bias rides in K columns 27 (hi) and 28 (lo)
  for (int i = threadIdx.x; i < NH * 4; i += BF_THREADS) {
    const int r = i >> 2, j = i & 3;
    const int o = s * NH + r;
    uint4 w = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(p.w1) + o * 128 + j * 16);
    if (j == 3) {
      const float b = p.b1[o];
      const uint32_t hi = pack_h2<F16>(b, 0.f) & 0xffffu;
      const float bhi = unpack_h2<F16>(hi).x;
      const uint32_t lo = pack_h2<F16>(b - bhi, 0.f) & 0xffffu;
      w.y = (w.y & 0x0000ffffu) | (hi << 16);           // k = 27
      w.z = (w.z & 0xffff0000u) | lo;                   // k = 28
    }
    *reinterpret_cast<uint4*>(smem_gen + BF_OFF_W1 + j * (NH * 16) + r * 16) = w;
  }
  fence_proxy_async_smem();

  if (warp == 0 && lane == 0) tma_prefetch_desc(&map_w2);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < BF_W_MAX; ++i) {
      mbar_init(smem_u32(&bar_wfull[i]), 1);
      mbar_init(smem_u32(&bar_wempty[i]), 1);
    }
    for (int i = 0; i < 6; ++i) {
      mbar_init(smem_u32(&bar_a1_full[i]), ALT ? 8 : 16);      // the T warps that build a batch (one set | both) x 2 CTAs
      mbar_init(smem_u32(&bar_c1_full[i]), 1);
      mbar_init(smem_u32(&bar_c1_empty[i]), ALT ? 8 : 16);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar_t_ready[i]), 16);
      mbar_init(smem_u32(&bar_t_free[i]), 1);
      mbar_init(smem_u32(&bar_acc_full[i]), 1);
      mbar_init(smem_u32(&bar_acc_empty[i]), 8);        // 4 epilogue warps x 2 CTAs
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&bar_sc_full[i]), EPI_PATCH ? 8 : 16);   // the warps that build it (4 epilogue | 8 T) x 2 CTAs
      mbar_init(smem_u32(&bar_sc_free[i]), 1);
      mbar_init(smem_u32(&bar_patch_ready[i]), 4);
      mbar_init(smem_u32(&bar_patch_free[i]), 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair(smem_u32(&tmem_base_slot), 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const long long cluster_id = blockIdx.x >> 1;
  const long long n_clusters = gridDim.x >> 1;
  const long long my_tiles = cluster_id < p.n_tiles ? (p.n_tiles - cluster_id + n_clusters - 1) / n_clusters : 0;
  const uint32_t t_base = smem_base + BF_OFF_T;
  const bool t_warp = (warp >= 4 && warp < 8) || warp >= 12;

  if (warp == 0) {
    // ================= weight producer: this CTA's 64 rows of every [128 x 64] K chunk, in consumption order =================
    if (elect_one()) {
      int ws = 0;
      uint32_t phase = 0;
      for (long long l = 0; l < my_tiles; ++l) {
        for (int ci = -1; ci < 2 * B::CHUNKS; ++ci) {   // the shortcut chunk, then phase 0's and phase 1's stages
          {
            int col = B::SC_COL;                        // first K column of the stage in the packed weights
            if (ci >= 0) {
              if (QUAD) {                               // taps (ky, kx0) and (ky, kx0 + 1): 128 consecutive columns
                const int ph = ci / B::CHUNKS, idx = ci % B::CHUNKS;
                col = (((ph == 0 ? 1 : 0) + 2 * (idx >> 1)) * 4 + 2 * (idx & 1)) * 64;
              } else {
                int ky, kx, h;
                bf_chunk(ci >> 4, ci & 15, ky, kx, h);
                col = ((ky * 4 + kx) * 2 + h) * 64;
              }
            }
            const bool two = QUAD && ci >= 0;           // two [NH x 64] boxes per stage; the shortcut chunk of CH = 64 is one
            const uint32_t bytes = (QUAD && ci < 0) ? BF_W_BYTES / 2 : BF_W_BYTES;
            mbar_wait_relaxed(smem_u32(&bar_wempty[ws]), phase ^ 1u);
            if (dbg & 32) {                             // timing experiment (wrong results): no weight loads at all
              if (leader) mbar_arrive(smem_u32(&bar_wfull[ws]));
            } else {
              const uint32_t wfull = mapa_u32(smem_u32(&bar_wfull[ws]), 0);
              if (leader) mbar_expect_tx(smem_u32(&bar_wfull[ws]), 2 * bytes);
              tma_load_2d_pair(smem_base + ws * BF_W_BYTES, &map_w2, wfull, col, (int)rank * NH);
              if (two) tma_load_2d_pair(smem_base + ws * BF_W_BYTES + BF_W_BYTES / 2, &map_w2, wfull, col + 64, (int)rank * NH);
            }
            if (++ws == w_stages) { ws = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= c2 MMA issuer (leader CTA only): 16 K chunks x 4 MMAs per unit =================
    // One thread's barrier instructions cost ~100 cycles each even when the phase is already complete, and a chunk is only 256
    // MMA cycles: the loop below touches ONE barrier per chunk and tests the next stage's barrier before issuing this stage's
    // MMAs so that the test's latency hides behind them; c1's batches have their own issuer (warp 3).
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc(256, CH, F16);
      // every address below is loop invariant; a single thread pays ~10 cycles per dependent instruction, so nothing is
      // recomputed per chunk: barrier addresses advance by 8, the weight descriptor by 8 KB >> 4, and T's descriptors are
      // compile-time offsets from t_base (the 16 chunks of a phase are unrolled)
      const uint32_t wfull0 = smem_u32(&bar_wfull[0]), wempty0 = smem_u32(&bar_wempty[0]);
      const uint32_t w_wrap = (uint32_t)w_stages * 8u;
      const uint64_t t_desc0 = make_nosw_desc(t_base, BF_K8, BF_ROW);
      const uint64_t w_desc0 = make_sw128_desc(smem_base);
      uint32_t wo = 0;                                  // 8 * stage
      uint32_t wphase = 0;
      bool ready = false;                               // stage wo / 8 is known to be full
      const uint64_t sc_desc = make_nosw_desc(smem_base + BF_OFF_SC, 2048, 128);
      for (long long l = 0; l < my_tiles; ++l) {
        const int acc = (int)(l & 1);
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * CH);
        {
          // the tile's first MMA initialises the accumulator with bias + W_sc . avg_pool2d(x): one K = 16 step whose operands
          // carry the fp32 terms as 16-bit hi / lo pairs (columns: ph.wh, pl.wh, ph.wl, 1.bias_hi, 1.bias_lo, 1.bias_lo2)
          const int sb = EPI_PATCH ? (int)(l & 1) : 0;                      // shortcut buffer of the tile and the parity of its use
          const uint32_t sc_par = EPI_PATCH ? (uint32_t)((l >> 1) & 1) : (uint32_t)(l & 1);
          mbar_wait(smem_u32(&bar_sc_full[sb]), sc_par);
          mbar_wait(smem_u32(&bar_acc_empty[acc]), (uint32_t)(((l >> 1) & 1) ^ 1));
          if (!ready) mbar_wait(wfull0 + wo, wphase);
          tc_fence_after();
          const uint32_t cur = wo;
          wo += 8u;
          if (wo == w_wrap) { wo = 0; wphase ^= 1u; }
          ready = mbar_try_wait(wfull0 + wo, wphase);
          umma_pair(d_tmem, sc_desc + (uint64_t)(sb * (BF_SC_BYTES >> 4)), w_desc0 + (uint64_t)(cur * (BF_W_BYTES / 16 / 8)), idesc, 0u);
          umma_commit_pair(wempty0 + cur, 3);
          umma_commit_pair(smem_u32(&bar_sc_free[sb]), 3);
        }
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
          mbar_wait(smem_u32(&bar_t_ready[ph]), (uint32_t)(l & 1));
          tc_fence_after();
#pragma unroll
          for (int idx = 0; idx < B::CHUNKS; ++idx) {
            constexpr int kStep = (2 * BF_K8) >> 4;
            if (!ready) mbar_wait(wfull0 + wo, wphase);
            tc_fence_after();
            const uint32_t cur = wo;
            wo += 8u;
            if (wo == w_wrap) { wo = 0; wphase ^= 1u; }
            ready = mbar_try_wait(wfull0 + wo, wphase);                     // next stage: consumed after the MMAs below
            const uint64_t bdesc = w_desc0 + (uint64_t)(cur * (BF_W_BYTES / 16 / 8));
            if (QUAD) {
              // stage idx = taps (ky, kx0) and (ky, kx0 + 1), 64 channels each: 2 x 4 MMAs of K = 16
              const int ky = (ph == 0 ? 1 : 0) + 2 * (idx >> 1), kx0 = 2 * (idx & 1);
              if (!(dbg & 16)) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int kx = kx0 + e;
                  const int a_off = ((((ky & 1) * 2 + (kx & 1)) * G) * BF_K8 + (ky >> 1) * BF_ROW + (kx >> 1) * 16) >> 4;
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    umma_pair(d_tmem, t_desc0 + (uint64_t)(a_off + j * kStep), bdesc + (uint64_t)(e * (BF_W_BYTES / 2 / 16) + 2 * j), idesc, 1u);
                }
              }
            } else {
              const int ky = (ph == 0 ? 1 : 0) + 2 * (idx >> 3), kx = (idx >> 1) & 3, h = idx & 1;      // = bf_chunk(ph, idx)
              const int a_off = ((((ky & 1) * 2 + (kx & 1)) * 16 + h * 8) * BF_K8 + (ky >> 1) * BF_ROW + (kx >> 1) * 16) >> 4;
              const uint64_t adesc = t_desc0 + (uint64_t)a_off;
              if (!(dbg & 16)) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  umma_pair(d_tmem, adesc + (uint64_t)(j * kStep), bdesc + (uint64_t)(2 * j), idesc, 1u);
              }
            }
            umma_commit_pair(wempty0 + cur, 3);
          }
          umma_commit_pair(smem_u32(&bar_t_free[ph]), 3);            // this half of T may be rebuilt once these MMAs retire
          if (ph == 1) umma_commit_pair(smem_u32(&bar_acc_full[acc]), 3);
        }
      }
    }
  } else if (warp == 3) {
    // ================= c1 MMA issuer (leader CTA only): one K = 32 batch of 256 pixels whenever both CTAs have built it =================
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc(256, CH, F16);
      const uint64_t a1desc = make_nosw_desc(smem_base + BF_OFF_A1, 2048, 128);
      const uint64_t w1desc = make_nosw_desc(smem_base + BF_OFF_W1, NH * 16, 128);
      const long long g_total = 6 * my_tiles;
      int ring = 0;                                     // ALT: position (g >> 1) % 3 in a set's ring of three buffers, and the
      uint32_t rpar = 0;                                // parity of its use, ((g >> 1) / 3) & 1
      for (long long g = 0; g < g_total; ++g) {
        // buffer / accumulator of batch g and the parity of its use: double-buffered, or (ALT) three per T-warp set (set = g & 1)
        const int cb = ALT ? (int)(g & 1) * 3 + ring : (int)(g & 1);
        const uint32_t par = ALT ? rpar : (uint32_t)((g >> 1) & 1);
        if (ALT && (g & 1)) { if (++ring == 3) { ring = 0; rpar ^= 1u; } }
        mbar_wait(smem_u32(&bar_a1_full[cb]), par);
        mbar_wait(smem_u32(&bar_c1_empty[cb]), par ^ 1u);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(2 * CH) + (uint32_t)(cb * CH);      // after c2's two accumulators
        const uint64_t ad = a1desc + (uint64_t)(cb * (BF_A1_BYTES >> 4));
        if (!(dbg & 8)) {
          umma_pair(d, ad, w1desc, idesc, 0u);                              // k = 0..15
          umma_pair(d, ad + 256u, w1desc + (uint64_t)(2 * NH), idesc, 1u);  // k = 16..31 (27, 28: bias; 29..31 zero)
        }
        umma_commit_pair(smem_u32(&bar_c1_full[cb]), 3);
      }
    }
  } else if (t_warp) {
    // ================= T warps: raw rows -> im2col batch -> [c1 MMA] -> relu / convert -> T cells =================
    // two sets of four warps (TMEM lane quadrant = warp & 3 in both): set 0 gathers K columns 0..15 and drains channels
    // 0..63 of every batch, set 1 the other halves
    const int set = warp >= 12 ? 1 : 0;
    const int q = warp & 3;
    const int tt = q * 32 + lane;                       // batch pixel = TMEM lane
    const int t256 = set * 128 + tt;
    constexpr uint32_t kOne = F16 ? 0x3C00u : 0x3F80u;
    uint32_t vmaxw = 0;                                 // fp16 range guard: running maximum of the (non-negative) packed halves
    if constexpr (ALT) {
      // CH = 64.  The two sets take ALTERNATE batches (set = g & 1), each with two im2col buffers / c1 accumulators of its own.
      // With a 64-channel c2 the tensor pipe needs a tile's T in ~2 400 cycles and these eight warps pace the kernel: two in-order
      // warps per scheduler issue one dependent instruction every ~5 cycles, so what counts is the NUMBER of instructions per
      // batch -- and the SIZE of the loop: a version specialised per set and batch at compile time was 40 % slower (32 KB of
      // T-warp code against a 6 KB L0 / 32 KB L1.5 instruction cache shared with the other roles).  So: ONE copy of the loop body,
      // and everything that does not depend on the data hoisted out of it.  A thread's three pixels of a tile (batch
      // r = set + 2 k: unit uu = r / 3, batch b = r % 3 of the unit) are described by one packed word each, computed once per
      // launch: patch offset, T cell, "outside the image" bit per quadrant, and the roles of the batch in the barrier protocol;
      // barrier addresses (incl. the leader's, mapa) are computed once, buffers and parities follow one counter
      // (kc = 3 l + k = g >> 1).
      constexpr uint32_t W_PC = 1u << 27, W_SKIP = 1u << 28, W_FIRST = 1u << 29, W_LAST = 1u << 30, W_UNIT1 = 1u << 31;
      constexpr uint32_t W_INVALID = 0x3ffu;            // bits 0-9 all ones: no pixel
      uint32_t pw[3];     // bits 0-9: patch offset / 8, 10-22: T cell offset / 16, 23-26: outside-the-image bit of quadrant 0..3
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int r = set + 2 * k, uu = r / 3, b = r - 3 * uu, pr = 1 - uu;
        int R, C, pc;
        const bool valid = bf_pixel<true>(s, pr, b, tt, R, C, pc);
        const int y = 2 * R + pr - 1, x = 16 * s - 1 + 2 * C + pc;        // pixel of the quadrant: rows -1 .. 32, columns 16 s - 1 .. 16 s + 16
        // tap (0, 0) of the pixel in the patch: patch row y + 1, patch column x - 1 - (16 s - 2) = 2 C + pc -> plane pc, entry C
        const int g_off = pc * BF_PQ_PLANE + (y + 1) * BF_PQ_ROWB + C * 8;
        const int c_off = ((pr * 2 + pc) * G) * BF_K8 + R * BF_ROW + C * 16;     // its T cell (channel group 0)
        uint32_t w = valid ? ((uint32_t)(g_off >> 3) | ((uint32_t)(c_off >> 4) << 10) | (pc ? W_PC : 0u)) : W_INVALID;
#pragma unroll
        for (int quad = 0; quad < 4; ++quad) {
          const int Y = 32 * (quad >> 1) + y, X = 32 * (quad & 1) + x;
          if (Y < 0 || Y >= 64 || X < 0 || X >= 64) w |= 1u << (23 + quad);
        }
        if (b == 2 && q >= 2) w |= W_SKIP;              // the third batch of a unit holds 50 pixels: lane quadrants 0, 1
        if (r == set || r == 4 - set) w |= W_FIRST;     // this set's first batch of the unit (r = 0, 4 | 1, 3)
        if (r == 2 + 3 * set || r == 4 - 3 * set) w |= W_LAST;   // ... and its last one (r = 2, 4 | 1, 5)
        if (uu) w |= W_UNIT1;
        pw[k] = w;
      }
      const uint32_t b_c1full = smem_u32(&bar_c1_full[0]), b_tfree = smem_u32(&bar_t_free[0]);
      const uint32_t b_pready = smem_u32(&bar_patch_ready[0]), b_pfree = smem_u32(&bar_patch_free[0]);
      const uint32_t r_a1full = mapa_u32(smem_u32(&bar_a1_full[0]), 0), r_c1empty = mapa_u32(smem_u32(&bar_c1_empty[0]), 0);
      const uint32_t r_tready = mapa_u32(smem_u32(&bar_t_ready[0]), 0);
      uint8_t* const a1_row = smem_gen + BF_OFF_A1 + tt * 16;
      const uint32_t t_lane = tmem_base + (uint32_t)(2 * CH) + ((uint32_t)(q * 32) << 16);

      const int n_my = (int)my_tiles;                   // (a cluster sees < 2^28 tiles)
      // A set's batches use a ring of THREE im2col buffers / c1 accumulators (set * 3 + 0..2).  While batch j = (l, k) is drained,
      // the operand of batch j + 2 is gathered: the round trip gather -> both CTAs' arrives -> c1 issuer -> MMA -> commit has
      // two iterations to complete (with two buffers and one iteration of slack the wait for it was 15 % of these warps' time),
      // and that slack also pays for signalling "operand built" together with "accumulator drained" after ONE proxy fence per
      // iteration instead of two.
      int db = 0;                                       // ring position of the batch being drained,
      uint32_t dpar = 0;                                // the parity of that use,
      int bb = 0;                                       // ring position the next gather writes
      long long t = cluster_id;
      // (l, k) = (-1, 1), (-1, 2): the prologue, only the gathers of tile 0's first two batches
      for (int l = -1; l < n_my; ++l) {
        const uint32_t tf_par = (uint32_t)((l & 1) ^ 1);
#pragma unroll 1
        for (int k = l < 0 ? 1 : 0; k < 3; ++k) {
          // ---- the operand of the batch after next: (l, 2) when k = 0, else (l + 1, k - 1); gather the 32 K columns of this
          // thread's pixel from the patch (its buffer was read by the batch before this one, whose completion this thread has seen)
          uint32_t wn = k == 0 ? pw[2] : (k == 1 ? pw[0] : pw[1]);
          const int ln = k == 0 ? l : l + 1;            // tile of that batch, patch buffer ln & 1
          if (ln >= n_my) wn = 0xffffffffu;             // no such tile: no gather, no arrive
          else if (k == 1) mbar_wait(b_pready + 8u * (uint32_t)(ln & 1), (uint32_t)((ln >> 1) & 1));    // its patch (epilogue warps)
          const int cbn = set * 3 + bb;
          if (wn != 0xffffffffu) {
            if ((wn & 0x3ffu) != W_INVALID && !(dbg & 1)) {
              // columns x - 1 and x + 1 in the pixel's own column-parity plane, column x in the other one
              const uint8_t* base = smem_gen + BF_OFF_P + (ln & 1) * BF_P_BYTES + ((wn & 0x3ffu) << 3);
              const uint8_t* mid = base + ((wn & W_PC) ? 8 - BF_PQ_PLANE : BF_PQ_PLANE);
              uint8_t* row = a1_row + cbn * BF_A1_BYTES;
              auto ld = [](const uint8_t* q_) { return *reinterpret_cast<const uint2*>(q_); };
              const uint2 a0 = ld(base), a1 = ld(mid), a2 = ld(base + 8);
              const uint2 b0 = ld(base + BF_PQ_ROWB), b1 = ld(mid + BF_PQ_ROWB), b2 = ld(base + BF_PQ_ROWB + 8);
              const uint2 c0 = ld(base + 2 * BF_PQ_ROWB), c1 = ld(mid + 2 * BF_PQ_ROWB), c2 = ld(base + 2 * BF_PQ_ROWB + 8);
              uint4 w;
              w.x = a0.x;
              w.y = __byte_perm(a0.y, a1.x, 0x5410);    // lo16(a0.y) | lo16(a1.x) << 16
              w.z = __byte_perm(a1.x, a1.y, 0x5432);    // hi16(a1.x) | lo16(a1.y) << 16
              w.w = a2.x;
              *reinterpret_cast<uint4*>(row) = w;
              w.x = __byte_perm(a2.y, b0.x, 0x5410);
              w.y = __byte_perm(b0.x, b0.y, 0x5432);
              w.z = b1.x;
              w.w = __byte_perm(b1.y, b2.x, 0x5410);
              *reinterpret_cast<uint4*>(row + 2048) = w;
              w.x = __byte_perm(b2.x, b2.y, 0x5432);
              w.y = c0.x;
              w.z = __byte_perm(c0.y, c1.x, 0x5410);
              w.w = __byte_perm(c1.x, c1.y, 0x5432);
              *reinterpret_cast<uint4*>(row + 2 * 2048) = w;
              w.x = c2.x;
              w.y = (c2.y & 0xffffu) | (kOne << 16);     // k = 27, 28: 1.0 (bias hi / lo)
              w.z = kOne;
              w.w = 0u;
              *reinterpret_cast<uint4*>(row + 3 * 2048) = w;
            }
            if (++bb == 3) bb = 0;
          }
          // ---- this batch: TMEM accumulator cb -> relu / convert -> T cells (all 64 channels) ----
          const uint32_t w = k == 0 ? pw[0] : (k == 1 ? pw[1] : pw[2]);
          const int cb = set * 3 + db;
          if (l >= 0) {
            mbar_wait(b_c1full + 8u * (uint32_t)cb, dpar);                  // the batch is in TMEM
            tc_fence_after();
            if (w & W_FIRST) mbar_wait(b_tfree + ((w & W_UNIT1) ? 8u : 0u), tf_par);   // c2 is done with this half of T
            if (!(w & W_SKIP) && !(dbg & 2)) {
              uint32_t r[64];
              tmem_ld32(t_lane + (uint32_t)(cb * CH), r);
              tmem_ld32(t_lane + (uint32_t)(cb * CH) + 32u, r + 32);
              tmem_ld_wait();
              if ((w & 0x3ffu) != W_INVALID) {
                const bool in_img = !((w >> (23 + (int)(t & 3))) & 1u);     // a cell outside the image is conv padding = zero
                uint8_t* cell = smem_gen + BF_OFF_T + (((w >> 10) & 0x1fffu) << 4);
#pragma unroll
                for (int gq = 0; gq < 8; ++gq) {
                  uint4 pk = make_uint4(0u, 0u, 0u, 0u);
                  if (in_img) {
                    pk.x = pack_relu_h2<F16>(__uint_as_float(r[gq * 8 + 0]), __uint_as_float(r[gq * 8 + 1]));
                    pk.y = pack_relu_h2<F16>(__uint_as_float(r[gq * 8 + 2]), __uint_as_float(r[gq * 8 + 3]));
                    pk.z = pack_relu_h2<F16>(__uint_as_float(r[gq * 8 + 4]), __uint_as_float(r[gq * 8 + 5]));
                    pk.w = pack_relu_h2<F16>(__uint_as_float(r[gq * 8 + 6]), __uint_as_float(r[gq * 8 + 7]));
                  }
                  if (F16) vmaxw = __vmaxu2(__vmaxu2(vmaxw, pk.x), __vmaxu2(__vmaxu2(pk.y, pk.z), pk.w));
                  *reinterpret_cast<uint4*>(cell + gq * BF_K8) = pk;
                }
                if (p.dbg_t && in_img) {                // tests only: the same 16-bit values to global memory
                  const int rr = set + 2 * k, uu = rr / 3;
                  int R, C, pc;
                  bf_pixel<true>(s, 1 - uu, rr - 3 * uu, tt, R, C, pc);
                  const int Y = 32 * ((int)(t & 3) >> 1) + 2 * R - uu, X = 32 * (int)(t & 1) + 16 * s - 1 + 2 * C + pc;
#pragma unroll 1
                  for (int gq = 0; gq < 8; ++gq)
                    *reinterpret_cast<uint4*>(p.dbg_t + ((((t >> 2) * IMG + Y) * IMG + X) * CH + gq * 8)) =
                        *reinterpret_cast<const uint4*>(cell + gq * BF_K8);
                }
              }
            }
            tc_fence_before();
          }
          // ---- one fence for the im2col rows and the T cells (both are read by the tensor core), then every signal ----
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (wn != 0xffffffffu) {
              mbar_arrive_cluster(r_a1full + 8u * (uint32_t)cbn);
              // after its last gather of a tile (batch (l, 2), gathered at k = 0) this warp is done with the patch buffer
              if (k == 0) mbar_arrive(b_pfree + 8u * (uint32_t)(l & 1));
            }
            if (l >= 0) {
              mbar_arrive_cluster(r_c1empty + 8u * (uint32_t)cb);
              // the unit is complete once every T warp of both CTAs has said so
              if (w & W_LAST) mbar_arrive_cluster(r_tready + ((w & W_UNIT1) ? 8u : 0u));
            }
          }
          if (l >= 0 && ++db == 3) { db = 0; dpar ^= 1u; }
        }
        if (l >= 0) t += n_clusters;
      }
    } else {
      // CH = 128: tile l of this cluster = image n; the landing buffer holds, per image row, the 56 bytes of image columns
      // 15 s - 1 .. 15 s + 17 from byte goff of the row (4-byte aligned), at offset 4 of a 64-byte landing row
      constexpr int XO = 4;
      const int goff = s ? 40 : 0;
      auto prefetch_x = [&](long long l) {              // raw rows of tile l -> the (single) landing buffer
        if (l < my_tiles) {
          const uint8_t* src = p.x + (cluster_id + l * n_clusters) * 3072 + goff;
          const uint32_t dst = smem_base + BF_OFF_X + XO;
          for (int w = t256; w < 32 * 14; w += BF_T_THREADS) {
            const int row = w / 14, wi = w - row * 14;
            cp_async_4(dst + row * BF_X_ROWB + wi * 4, src + row * 96 + wi * 4);
          }
        }
        cp_async_commit();
      };
      auto nrm = [](uint8_t u) { return bf_nrm(u); };
      // landed bytes -> (a) normalised 16-bit patch, the first conv's operand values, converted ONCE per byte instead of once per
      // tap: patch pixel (ry, cx) = image pixel (ry - 1, 15 s - 1 + cx) as (c0, c1, c2, 0), 18 in-image columns per row;
      // (b) the shortcut operand of tile L: row m = avg_pool2d(normalised x) at output pixel (m >> 3, 8 s + (m & 7)) as hi / lo pairs
      auto convert_x = [&](long long L) {
        const uint8_t* raw = smem_gen + BF_OFF_X + XO;
        for (int i = t256; i < 32 * 18; i += BF_T_THREADS) {
          const int row = i / 18, j = i - row * 18;     // j-th in-image pixel of the row: image column (s ? 14 : 0) + j
          const uint8_t* b = raw + row * BF_X_ROWB + 3 * ((s ? 14 : 0) + j) - goff;
          const uint32_t lo = pack_h2<F16>(nrm(b[0]), nrm(b[1])), hi = pack_h2<F16>(nrm(b[2]), 0.f);
          *reinterpret_cast<uint2*>(smem_gen + BF_OFF_P + (row + 1) * BF_P_ROWB + (j + (s ? 0 : 1)) * 8) = make_uint2(lo, hi);
        }
        mbar_wait(smem_u32(&bar_sc_free[0]), (uint32_t)((L & 1) ^ 1));  // the previous tile's shortcut MMA has read the buffer
        {
          // the 2 x 2 input pixels under pooled output pixel (tt >> 3, 8 s + (tt & 7)) of the tile
          const uint8_t* b = raw + (2 * (tt >> 3)) * BF_X_ROWB + 3 * (2 * (8 * s + (tt & 7))) - goff;
          float px[3];
#pragma unroll
          for (int c = 0; c < 3; ++c)
            px[c] = (nrm(b[c]) + nrm(b[3 + c]) + nrm(b[BF_X_ROWB + c]) + nrm(b[BF_X_ROWB + 3 + c])) * 0.25f;
          uint32_t ph[3], pl[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            ph[c] = pack_h2<F16>(px[c], 0.f) & 0xffffu;
            pl[c] = pack_h2<F16>(px[c] - unpack_h2<F16>(ph[c]).x, 0.f) & 0xffffu;
          }
          uint4 v;
          if (set == 0) v = make_uint4(ph[0] | (ph[1] << 16), ph[2] | (pl[0] << 16), pl[1] | (pl[2] << 16), ph[0] | (ph[1] << 16));
          else v = make_uint4(ph[2] | (kOne << 16), kOne | (kOne << 16), 0u, 0u);
          *reinterpret_cast<uint4*>(smem_gen + BF_OFF_SC + set * 2048 + tt * 16) = v;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bar_sc_full[0]), 0));
      };
      // gather this set's 16 K columns of batch pixel tt of (unit uu, batch b) of the tile from the patch into A1[gb]:
      // K column k = (ky*3 + kx)*3 + c for k < 27; 27, 28 = 1.0 (bias hi / lo); 29..31 = 0
      auto build_a1 = [&](int uu, int b, int gb) {
        const int pr = 1 - uu;
        int R, C, pc;
        const bool valid = bf_pixel<false>(s, pr, b, tt, R, C, pc);
        if (valid && !(dbg & 1)) {
          const int y = 2 * R + pr - 1, x = 16 * s - 1 + 2 * C + pc;     // pixel of the tile: rows -1 .. 32, columns 16 s - 1 .. 16 s + 16
          // tap (ky, kx) reads image pixel (y + ky - 1, x + kx - 1) = patch row y + ky, patch column x - 15 s + kx
          const uint2* row0 = reinterpret_cast<const uint2*>(smem_gen + BF_OFF_P + y * BF_P_ROWB + (x - 15 * s) * 8);
          const uint2* row1 = reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(row0) + BF_P_ROWB);
          const uint2* row2 = reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(row0) + 2 * BF_P_ROWB);
          uint8_t* row = smem_gen + BF_OFF_A1 + gb * BF_A1_BYTES + tt * 16;
          if (set == 0) {
            const uint2 a0 = row0[0], a1 = row0[1], a2 = row0[2], b0 = row1[0], b1 = row1[1], b2 = row1[2];
            uint32_t w[8];
            w[0] = a0.x;
            w[1] = __byte_perm(a0.y, a1.x, 0x5410);     // lo16(a0.y) | lo16(a1.x) << 16
            w[2] = __byte_perm(a1.x, a1.y, 0x5432);     // hi16(a1.x) | lo16(a1.y) << 16
            w[3] = a2.x;
            w[4] = __byte_perm(a2.y, b0.x, 0x5410);
            w[5] = __byte_perm(b0.x, b0.y, 0x5432);
            w[6] = b1.x;
            w[7] = __byte_perm(b1.y, b2.x, 0x5410);
            *reinterpret_cast<uint4*>(row) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(row + 2048) = make_uint4(w[4], w[5], w[6], w[7]);
          } else {
            const uint2 b2 = row1[2], c0 = row2[0], c1 = row2[1], c2 = row2[2];
            uint32_t w[8];
            w[0] = __byte_perm(b2.x, b2.y, 0x5432);
            w[1] = c0.x;
            w[2] = __byte_perm(c0.y, c1.x, 0x5410);
            w[3] = __byte_perm(c1.x, c1.y, 0x5432);
            w[4] = c2.x;
            w[5] = (c2.y & 0xffffu) | (kOne << 16);
            w[6] = kOne;
            w[7] = 0u;
            *reinterpret_cast<uint4*>(row + 2 * 2048) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(row + 3 * 2048) = make_uint4(w[4], w[5], w[6], w[7]);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bar_a1_full[gb]), 0));
      };
      // batch (uu, b) of image n is in TMEM accumulator cb: relu / convert -> T cells, this set's 64 channels
      auto drain = [&](long long n, int uu, int b, int cb) {
        constexpr int NCH = CH / 2;
        const int ch0 = set * (CH / 2);
        const int pr = 1 - uu;
        int R, C, pc;
        const bool valid = bf_pixel<false>(s, pr, b, tt, R, C, pc);
        if ((b < 2 || q < 1) && !(dbg & 2)) {            // the third batch of a unit holds 16 pixels: lane quadrant 0
          uint8_t* cell = smem_gen + BF_OFF_T + ((pr * 2 + pc) * G + (ch0 >> 3)) * BF_K8 + R * BF_ROW + C * 16;
          const int Y = 2 * R + pr - 1, X = 16 * s - 1 + 2 * C + pc;     // pixel of the image (the halo cells are never computed)
          const uint32_t taddr = tmem_base + (uint32_t)(2 * CH) + (uint32_t)(cb * CH + ch0) + ((uint32_t)(q * 32) << 16);
          uint32_t r[NCH];
          tmem_ld32(taddr, r);
          tmem_ld32(taddr + 32u, r + 32);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int gq = 0; gq < NCH / 8; ++gq) {
              uint4 pk;
              pk.x = pack_relu_h2<F16>(__uint_as_float(r[gq * 8 + 0]), __uint_as_float(r[gq * 8 + 1]));
              pk.y = pack_relu_h2<F16>(__uint_as_float(r[gq * 8 + 2]), __uint_as_float(r[gq * 8 + 3]));
              pk.z = pack_relu_h2<F16>(__uint_as_float(r[gq * 8 + 4]), __uint_as_float(r[gq * 8 + 5]));
              pk.w = pack_relu_h2<F16>(__uint_as_float(r[gq * 8 + 6]), __uint_as_float(r[gq * 8 + 7]));
              if (F16) vmaxw = __vmaxu2(__vmaxu2(vmaxw, pk.x), __vmaxu2(__vmaxu2(pk.y, pk.z), pk.w));
              *reinterpret_cast<uint4*>(cell + gq * BF_K8) = pk;
              if (p.dbg_t)
                *reinterpret_cast<uint4*>(p.dbg_t + ((((long long)n * IMG + Y) * IMG + X) * CH + ch0 + gq * 8)) = pk;
            }
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();                       // T cells -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bar_c1_empty[cb]), 0));
      };
      // the patch of tile l replaces the previous one: every T thread has landed its rows and nobody reads the old patch any more
      auto switch_patch = [&](long long l) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        named_bar_sync(1, BF_T_THREADS);
        convert_x(l);
        named_bar_sync(1, BF_T_THREADS);                // patch complete, landing buffer idle
        prefetch_x(l + 1);
      };
      if (my_tiles > 0) {
        prefetch_x(0);
        switch_patch(0);
        build_a1(0, 0, 0);
      }
      long long g = 0;
      for (long long l = 0; l < my_tiles; ++l) {
        const long long n = cluster_id + l * n_clusters;
        for (int uu = 0; uu < 2; ++uu) {
          for (int b = 0; b < 3; ++b, ++g) {
            const int cb = (int)(g & 1);
            // the next batch's operand first (its buffer was read by batch g - 1, whose completion this thread has seen), so
            // that its MMA overlaps this batch's drain
            if (b < 2) build_a1(uu, b + 1, cb ^ 1);
            else if (uu == 0) build_a1(1, 0, cb ^ 1);
            else if (l + 1 < my_tiles) {
              switch_patch(l + 1);
              build_a1(0, 0, cb ^ 1);
            }
            mbar_wait(smem_u32(&bar_c1_full[cb]), (uint32_t)((g >> 1) & 1));      // batch g is in TMEM
            tc_fence_after();
            if (b == 0) mbar_wait(smem_u32(&bar_t_free[uu]), (uint32_t)((l & 1) ^ 1));   // c2 is done with this half of T
            drain(n, uu, b, cb);
          }
          if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bar_t_ready[uu]), 0));
        }
      }
    }
    // post-ReLU halves are non-negative: inf / NaN <=> a half >= 0x7C00
    if (F16 && ((vmaxw & 0xffffu) >= 0x7C00u || (vmaxw >> 16) >= 0x7C00u)) range_flag_set(p.ovf, SDG_RANGE_ACT);
  } else if (warp >= 8 && warp < 12) {
    // ================= c2 epilogue: this CTA's 128 output pixels x 128 channels =================
    const int q = warp & 3;
    const int g4 = lane >> 2, i4 = lane & 3;
    uint32_t vmaxw = 0;
    // ---- EPI_PATCH (CH = 64): these warps also stage the inputs of the T path, two tiles ahead of their own epilogue ----
    // raw rows of tile L -> landing buffer L & 1 (cp.async); then, one iteration later, -> (a) the normalised 16-bit patch L & 1:
    // patch pixel (ry, cx) = quadrant pixel (ry - 2, 16 s - 2 + cx) as (c0, c1, c2, 0), zeros outside the IMAGE; (b) the shortcut
    // operand L & 1: row m = avg_pool2d(normalised x) at output pixel (m >> 3, 8 s + (m & 7)) as hi / lo 16-bit pairs (K = 16: pooled
    // pixel hi x 3, lo x 3, hi x 3 again, 1, 1, 1 against W_sc hi, hi, lo and the bias in three pieces; see b1_fused_pack_kernel)
    const int e128 = q * 32 + lane;
    auto e_tile_of = [&](long long L, long long& n, int& qy, int& qx, int& goff) {
      const long long t = cluster_id + L * n_clusters;
      n = t >> 2; qy = (int)(t & 3) >> 1; qx = (int)(t & 1);
      const int b0 = (32 * qx + 16 * s - 2) * 3;
      goff = b0 < 0 ? 0 : (b0 & ~3);
    };
    auto e_prefetch = [&](long long L) {
      if (L < my_tiles) {
        long long n; int qy, qx, goff;
        e_tile_of(L, n, qy, qx, goff);
        const uint32_t dst = smem_base + BF_OFF_X + (uint32_t)(L & 1) * B::X_BYTES;
        const uint8_t* src = p.x + n * (64 * 64 * 3) + goff;
        for (int w = e128; w < 36 * 16; w += 128) {
          const int row = w >> 4, wi = w & 15;
          const int Y = 32 * qy - 2 + row;
          if (Y >= 0 && Y < 64 && goff + wi * 4 < 192) cp_async_4(dst + row * BF_X_ROWB + wi * 4, src + Y * 192 + wi * 4);
        }
      }
      cp_async_commit();
    };
    auto e_produce = [&](long long L) {
      if (L >= my_tiles) return;
      const int pb = (int)(L & 1);
      const uint32_t par = (uint32_t)(((L >> 1) & 1) ^ 1);
      const uint8_t* raw = smem_gen + BF_OFF_X + pb * B::X_BYTES;
      uint8_t* patch = smem_gen + BF_OFF_P + pb * BF_P_BYTES;
      long long n_; int qy, qx, goff;
      e_tile_of(L, n_, qy, qx, goff);
      // Pixel X of an image row sits at byte 3 X - goff of its landing row.  The patch starts at image column X0 = 32 qx + 16 s - 2
      // and 3 X0 = 2 (mod 4): patch column cx is at byte 3 cx + 2 (goff = 3 X0 - 2), or at 3 cx - 6 for the first strip of the image
      // (X0 = -2, goff = 0).  Either way four pixels are the 12 bytes from byte 2 of an aligned word: four LDS.32 per four pixels
      // instead of a byte load and a table lookup per value; the conversion is bf_nrm()'s arithmetic (these warps have the time,
      // the shared-memory pipe has not: profiles/r4_b1fused64.md).
      const int first = (qx | s) == 0 ? 8 : 0;
      mbar_wait(smem_u32(&bar_patch_free[pb]), par);    // the T warps have gathered tile L - 2 from this buffer
#pragma unroll 1
      for (int i = e128; i < 36 * 5; i += 128) {
        const int ry = i / 5, g = i - ry * 5;           // patch row, group of four patch columns 4 g .. 4 g + 3
        const int Y = 32 * qy - 2 + ry, X0 = 32 * qx + 16 * s - 2 + 4 * g;
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0u;
        if (Y >= 0 && Y < 64) {
          const int boff = 12 * g - first;              // byte offset of the first of the four words (< 0: pixels left of the image)
          const uint32_t* wr = reinterpret_cast<const uint32_t*>(raw + ry * BF_X_ROWB + (boff < 0 ? 0 : boff));
          uint32_t w0 = wr[0], w1 = wr[1], w2 = wr[2], w3 = wr[3];
          if (boff < 0) { w2 = w0; w3 = w1; }           // boff = -8: words 2, 3 are the row's first two
          const uint32_t by[12] = {(w0 >> 16) & 255u, w0 >> 24, w1 & 255u, (w1 >> 8) & 255u, (w1 >> 16) & 255u, w1 >> 24,
                                   w2 & 255u, (w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24, w3 & 255u, (w3 >> 8) & 255u};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (X0 + j >= 0 && X0 + j < 64) {
              v[2 * j] = pack_h2<F16>(bf_nrm((uint8_t)by[3 * j]), bf_nrm((uint8_t)by[3 * j + 1]));
              v[2 * j + 1] = pack_h2<F16>(bf_nrm((uint8_t)by[3 * j + 2]), 0.f);
            }
          }
        }
        // patch columns 4 g, 4 g + 2 -> entries 2 g, 2 g + 1 of the even plane; 4 g + 1, 4 g + 3 -> the same entries of the odd one
        *reinterpret_cast<uint4*>(patch + ry * BF_PQ_ROWB + g * 16) = make_uint4(v[0], v[1], v[4], v[5]);
        *reinterpret_cast<uint4*>(patch + BF_PQ_PLANE + ry * BF_PQ_ROWB + g * 16) = make_uint4(v[2], v[3], v[6], v[7]);
      }
      mbar_wait(smem_u32(&bar_sc_free[pb]), par);       // the shortcut MMA of tile L - 2 has read this buffer
      {
        // the 2 x 2 input pixels under pooled output pixel (e128 >> 3, 8 s + (e128 & 7)) of the tile: image column
        // 32 qx + 2 (8 s + j) = X0 + 2 + 2 j, i.e. six bytes from byte 8 + 6 j (6 j for the first strip) of landing rows 2 + 2 oy, + 1
        constexpr uint32_t kOne = F16 ? 0x3C00u : 0x3F80u;
        const int off = 6 * (e128 & 7) + 8 - first;
        const uint32_t sh = (uint32_t)(off & 3) * 8u;
        const uint32_t* r0 = reinterpret_cast<const uint32_t*>(raw + (2 + 2 * (e128 >> 3)) * BF_X_ROWB + (off & ~3));
        const uint32_t* r1 = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(r0) + BF_X_ROWB);
        const uint32_t a0 = r0[0], a1 = r0[1], c0 = r1[0], c1 = r1[1];
        const uint32_t lo0 = __funnelshift_r(a0, a1, sh), hi0 = a1 >> sh, lo1 = __funnelshift_r(c0, c1, sh), hi1 = c1 >> sh;
        // bytes 0..2 = left pixel, 3..5 = right pixel of the row
        const uint32_t l0[3] = {lo0 & 255u, (lo0 >> 8) & 255u, (lo0 >> 16) & 255u}, q0[3] = {lo0 >> 24, hi0 & 255u, (hi0 >> 8) & 255u};
        const uint32_t l1[3] = {lo1 & 255u, (lo1 >> 8) & 255u, (lo1 >> 16) & 255u}, q1[3] = {lo1 >> 24, hi1 & 255u, (hi1 >> 8) & 255u};
        uint32_t ph[3], pl[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float px = (bf_nrm((uint8_t)l0[c]) + bf_nrm((uint8_t)q0[c]) + bf_nrm((uint8_t)l1[c]) + bf_nrm((uint8_t)q1[c])) * 0.25f;
          ph[c] = pack_h2<F16>(px, 0.f) & 0xffffu;
          pl[c] = pack_h2<F16>(px - unpack_h2<F16>(ph[c]).x, 0.f) & 0xffffu;
        }
        uint8_t* sc = smem_gen + BF_OFF_SC + pb * BF_SC_BYTES + e128 * 16;
        *reinterpret_cast<uint4*>(sc) = make_uint4(ph[0] | (ph[1] << 16), ph[2] | (pl[0] << 16), pl[1] | (pl[2] << 16), ph[0] | (ph[1] << 16));
        *reinterpret_cast<uint4*>(sc + 2048) = make_uint4(ph[2] | (kOne << 16), kOne | (kOne << 16), 0u, 0u);
      }
      fence_proxy_async_smem();                         // the shortcut operand is read by the tensor core
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&bar_patch_ready[pb]));
        mbar_arrive_cluster(mapa_u32(smem_u32(&bar_sc_full[pb]), 0));
      }
    };
    if (EPI_PATCH) e_prefetch(0);
    for (long long l = EPI_PATCH ? -2 : 0; l < my_tiles; ++l) {
      if (EPI_PATCH) {
        // rows of tile l + 2 have landed (every thread's, after the barrier); the barrier also says that everybody is done reading
        // the other landing buffer (tile l + 1, converted in the previous iteration), which the next prefetch overwrites
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        named_bar_sync(2, 128);
        e_prefetch(l + 3);
        e_produce(l + 2);
        if (l < 0) continue;
      }
      const long long t = cluster_id + l * n_clusters;
      const long long n = QUAD ? t >> 2 : t;
      const int qy = QUAD ? (int)(t & 3) >> 1 : 0, qx = QUAD ? (int)(t & 1) : 0;
      const int acc = (int)(l & 1);
      mbar_wait_relaxed(smem_u32(&bar_acc_full[acc]), (uint32_t)((l >> 1) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(acc * CH) + ((uint32_t)(q * 32) << 16);
      if (!(dbg & 4)) {
#pragma unroll 1
      for (int c0 = 0; c0 < CH; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)c0, r);
        tmem_ld_wait();
        // bias and shortcut are already in the accumulator (the tile's first MMA): ReLU + convert only
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          pk[j] = pack_relu_h2<F16>(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
          if (F16) vmaxw = __vmaxu2(vmaxw, pk[j]);
        }
        // 4 x 4 transpose inside each group of 4 lanes: 4 lanes then write the 64 contiguous bytes of ONE pixel
        warp_transpose4_u4(pk, lane);
#pragma unroll
        for (int mm = 0; mm < 4; ++mm) {
          const int m2 = q * 32 + g4 * 4 + mm;
          const long long opix = (n * (IMG / 2) + 16 * qy + (m2 >> 3)) * (IMG / 2) + 16 * qx + 8 * s + (m2 & 7);
          *reinterpret_cast<uint4*>(p.out_relu + opix * CH + c0 + i4 * 8) = make_uint4(pk[4 * mm], pk[4 * mm + 1], pk[4 * mm + 2], pk[4 * mm + 3]);
        }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bar_acc_empty[acc]), 0));
    }
    if (F16 && ((vmaxw & 0xffffu) >= 0x7C00u || (vmaxw >> 16) >= 0x7C00u)) range_flag_set(p.ovf, SDG_RANGE_ACT);
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

int b1_fused_init() {
  SDG_CUDA(cudaFuncSetAttribute((b1_fused_kernel<true, 128>), cudaFuncAttributeMaxDynamicSharedMemorySize, Bf<128>::SMEM));
  SDG_CUDA(cudaFuncSetAttribute((b1_fused_kernel<false, 128>), cudaFuncAttributeMaxDynamicSharedMemorySize, Bf<128>::SMEM));
  SDG_CUDA(cudaFuncSetAttribute((b1_fused_kernel<true, 64>), cudaFuncAttributeMaxDynamicSharedMemorySize, Bf<64>::SMEM));
  SDG_CUDA(cudaFuncSetAttribute((b1_fused_kernel<false, 64>), cudaFuncAttributeMaxDynamicSharedMemorySize, Bf<64>::SMEM));
  return 0;
}

// w2f[o][0 .. 16 CH) = w2[o][.] (when w2 is given; else the caller packed them in place), then the shortcut chunk:
// columns 16 CH .. = wh0 wh1 wh2 | wh0 wh1 wh2 | wl0 wl1 wl2 | bias_hi bias_lo bias_lo2 | zeros, w = wh + wl the fp32 W_sc / sigma
template <bool F16>
__global__ void __launch_bounds__(256)
b1_fused_pack_kernel(const h16* __restrict__ w2, const float* __restrict__ sc_w3, const float* __restrict__ bias2,
                     h16* __restrict__ w2f, int ch) {
  const int o = blockIdx.x;
  const int kmain = 16 * ch, ld = kmain + 64;
  if (w2)
    for (int i = threadIdx.x; i < kmain / 8; i += blockDim.x)
      reinterpret_cast<uint4*>(w2f + (size_t)o * ld)[i] = reinterpret_cast<const uint4*>(w2 + (size_t)o * kmain)[i];
  if (threadIdx.x < 64) {
    const int k = threadIdx.x;
    auto hi16 = [](float v) { return (uint16_t)(pack_h2<F16>(v, 0.f) & 0xffffu); };
    auto back = [](uint16_t h) { return unpack_h2<F16>((uint32_t)h).x; };
    uint16_t v = 0;
    if (k < 9) {
      const float w = sc_w3[o * 3 + k % 3];
      const uint16_t wh = hi16(w);
      v = k < 6 ? wh : hi16(w - back(wh));
    } else if (k < 12) {
      const float b = bias2[o];
      const uint16_t b0 = hi16(b);
      const uint16_t b1 = hi16(b - back(b0));
      v = k == 9 ? b0 : (k == 10 ? b1 : hi16(b - back(b0) - back(b1)));
    }
    w2f[(size_t)o * ld + kmain + k] = v;
  }
}

// ch = 128 (SNGAN-32) or 64 (SNGAN-64)
int b1_fused_pack(const h16* w2, const float* sc_w3, const float* bias2, h16* w2f, int ch, int f16, cudaStream_t s) {
  SDG_REQUIRE(sc_w3 && bias2 && w2f, SDG_E_INVALID, "b1_fused_pack: null pointer");
  SDG_REQUIRE(ch == 128 || ch == 64, SDG_E_UNSUPPORTED, "b1_fused_pack: ch=%d", ch);
  SDG_REQUIRE(((uintptr_t)w2 % 16) == 0 && ((uintptr_t)w2f % 16) == 0, SDG_E_INVALID, "b1_fused_pack: misaligned pointer");
  if (f16) { SDG_LAUNCH(b1_fused_pack_kernel<true>, ch, 256, 0, s, w2, sc_w3, bias2, w2f, ch); }
  else { SDG_LAUNCH(b1_fused_pack_kernel<false>, ch, 256, 0, s, w2, sc_w3, bias2, w2f, ch); }
  return 0;
}

int b1_fused_w2_ld(int ch) { return 16 * ch + 64; }
int b1_fused_w2_elems(int ch) { return ch * b1_fused_w2_ld(ch); }

template <bool F16, int CH>
static int b1_fused_launch(const CUtensorMap& map_w2, BfParams& p, cudaStream_t s) {
  using B = Bf<CH>;
  static const int ws_env = getenv("SDG_B1_WSTAGES") ? atoi(getenv("SDG_B1_WSTAGES")) : B::W_STAGES;
  p.w_stages = ws_env >= 2 && ws_env <= B::W_STAGES ? ws_env : B::W_STAGES;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(tc_num_sms() / 2 * 2));
  cfg.blockDim = dim3(BF_THREADS);
  cfg.dynamicSmemBytes = B::SMEM;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  // the tile schedule is static (tile = cluster + l * clusters): every cluster of the grid must be resident at once, and a GPC
  // with an odd number of free SMs cannot host a CTA pair on its last one -- ask the driver how many pairs fit
  long long clusters = tc_max_active_clusters((const void*)b1_fused_kernel<F16, CH>, &cfg);
  if (clusters < 1) clusters = 1;
  if (p.n_tiles < clusters) clusters = p.n_tiles;
  cfg.gridDim = dim3((unsigned)(2 * clusters));
  SDG_CUDA(cudaLaunchKernelEx(&cfg, b1_fused_kernel<F16, CH>, map_w2, p));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

// ch = 128: x [n][32][32][3] -> out_relu [n][16][16][128]; ch = 64: x [n][64][64][3] -> out_relu [n][32][32][64]
int b1_fused(const void* x, const h16* w1, const float* b1, const h16* w2f, h16* out_relu, h16* dbg_t, int64_t n, int ch, int f16,
             cudaStream_t s) {
  SDG_REQUIRE(x && w1 && b1 && w2f && out_relu, SDG_E_INVALID, "b1_fused: null pointer");
  SDG_REQUIRE(ch == 128 || ch == 64, SDG_E_UNSUPPORTED, "b1_fused: ch=%d", ch);
  auto al16 = [](const void* q) { return ((uintptr_t)q % 16) == 0; };
  SDG_REQUIRE(((uintptr_t)x % 4) == 0 && al16(w1) && al16(w2f) && al16(out_relu) && al16(dbg_t), SDG_E_INVALID,
              "b1_fused: misaligned pointer");
  if (n == 0) return 0;
  CUtensorMap map_w2;
  { int rc = tc_encode_2d(&map_w2, w2f, f16, b1_fused_w2_ld(ch), ch, 64, ch / 2); if (rc) return rc; }
  BfParams p;
  p.x = (const uint8_t*)x; p.w1 = w1; p.b1 = b1; p.out_relu = out_relu; p.dbg_t = dbg_t;
  p.ovf = t_range_flag; p.n_tiles = ch == 64 ? 4 * n : n;
  p.dbg = 0;
#ifdef SDG_TIMING_EXPERIMENTS
  static const int dbg_env = getenv("SDG_B1_DEBUG") ? atoi(getenv("SDG_B1_DEBUG")) : 0;
  p.dbg = dbg_env;
#endif
  if (ch == 128) return f16 ? b1_fused_launch<true, 128>(map_w2, p, s) : b1_fused_launch<false, 128>(map_w2, p, s);
  return f16 ? b1_fused_launch<true, 64>(map_w2, p, s) : b1_fused_launch<false, 64>(map_w2, p, s);
}

}  // namespace sdg
