// select.cu -- top-k index selection with the tie-break of a stable argsort.
//
// Replaces np.argsort(w)[-k:] / [:k] (eval_gan_drs_with_index.py:97-99, eval_gan_with_index.py:93-95,
// plot.py:100-101).  A stable ascending argsort orders samples by the composite key (score, index),
// which is a strict total order; the last / first k of it are therefore found by a radix select over
// the 96-bit composite (64-bit order-preserving image of the double, 32-bit index).
//
// ONE launch (select_fused_kernel, cooperative so that every CTA is resident): up to nine digit passes (11 bits each:
// six over the key, three over the index), each a grid-wide histogram of the still-matching candidates, a grid barrier and
// a redundant in-CTA scan of the bins; the passes stop as soon as the bin holding the k-th element contains exactly the
// candidates still needed (after 3-4 passes for distinct scores; ties at the threshold walk on into the index digits).
// Then one compaction sweep, and the last CTA to finish sorts the k survivors (bitonic, shared memory).  HBM-bound for
// large n (8 B per sample per pass), latency-bound at the 50 k samples of BASELINE configs[1] (round 1: 15 launches,
// 139 us; this kernel: one launch).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace sdg {

constexpr int kTopkMax = 4096;

struct SelectState {
  unsigned int hist[256];
  unsigned long long prefix_key;
  unsigned int prefix_idx;
  unsigned int k_rem;
  unsigned int ticket;
  unsigned int out_count;
};

__device__ __forceinline__ unsigned long long ordered_key(double x) {
  unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}

// composite in "select largest" space: for the smallest-k query both halves are complemented
__device__ __forceinline__ void composite(double x, unsigned int idx, int largest, unsigned long long& key,
                                          unsigned int& id) {
  key = ordered_key(x);
  id = idx;
  if (!largest) { key = ~key; id = ~id; }
}

__global__ void select_init_kernel(SelectState* st, unsigned int k) {
  int i = threadIdx.x;
  if (i < 256) st->hist[i] = 0;
  if (i == 0) { st->prefix_key = 0; st->prefix_idx = 0; st->k_rem = k; st->ticket = 0; st->out_count = 0; }
}

__global__ void __launch_bounds__(256)
select_pass_kernel(const double* __restrict__ score, int64_t n, int largest, int pass, SelectState* st) {
  __shared__ unsigned int sh[256];
  __shared__ bool is_last;
  sh[threadIdx.x] = 0;
  __syncthreads();
  const unsigned long long pk = st->prefix_key;
  const unsigned int pi = st->prefix_idx;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    unsigned long long key; unsigned int id;
    composite(score[i], (unsigned int)i, largest, key, id);
    bool match; unsigned int digit;
    if (pass < 8) {
      int hs = 64 - 8 * pass;                       // bits already fixed: the top 8*pass
      match = (pass == 0) || ((key >> hs) == (pk >> hs));
      digit = (unsigned int)(key >> (56 - 8 * pass)) & 255u;
    } else {
      int q = pass - 8;
      int hs = 32 - 8 * q;
      match = (key == pk) && ((q == 0) || ((id >> hs) == (pi >> hs)));
      digit = (id >> (24 - 8 * q)) & 255u;
    }
    if (match) atomicAdd(&sh[digit], 1u);
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], sh[threadIdx.x]);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  // last block: walk the bins from the top, find the digit holding the k_rem-th largest candidate
  __threadfence();
  sh[threadIdx.x] = atomicAdd(&st->hist[threadIdx.x], 0u);     // coherent read of the global bins
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int k = st->k_rem, above = 0;
    int d = 255;
    for (; d > 0; --d) {
      if (above + sh[d] >= k) break;
      above += sh[d];
    }
    st->k_rem = k - above;
    if (pass < 8) st->prefix_key = pk | ((unsigned long long)d << (56 - 8 * pass));
    else st->prefix_idx = pi | ((unsigned int)d << (24 - 8 * (pass - 8)));
    st->ticket = 0;
  }
  st->hist[threadIdx.x] = 0;
}

__global__ void __launch_bounds__(256)
select_compact_kernel(const double* __restrict__ score, int64_t n, int largest, SelectState* st,
                      unsigned long long* __restrict__ sel_key, unsigned int* __restrict__ sel_idx, int k) {
  const unsigned long long tk = st->prefix_key;
  const unsigned int ti = st->prefix_idx;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    unsigned long long key; unsigned int id;
    composite(score[i], (unsigned int)i, largest, key, id);
    if (key > tk || (key == tk && id >= ti)) {
      unsigned int slot = atomicAdd(&st->out_count, 1u);
      if (slot < (unsigned int)k) {
        sel_key[slot] = ordered_key(score[i]);      // original (ascending) ordering for the final sort
        sel_idx[slot] = (unsigned int)i;
      }
    }
  }
}

// single block: bitonic sort of k (<= 4096) composites ascending, then write indices
__global__ void __launch_bounds__(1024)
select_sort_kernel(const unsigned long long* __restrict__ sel_key, const unsigned int* __restrict__ sel_idx,
                   int k, int kpow2, int64_t* __restrict__ idx_out) {
  extern __shared__ unsigned char smem_raw[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
  unsigned int* ids = reinterpret_cast<unsigned int*>(keys + kpow2);
  for (int i = threadIdx.x; i < kpow2; i += blockDim.x) {
    keys[i] = i < k ? sel_key[i] : ~0ULL;
    ids[i] = i < k ? sel_idx[i] : ~0u;
  }
  __syncthreads();
  for (int size = 2; size <= kpow2; size <<= 1) {
    for (int str = size >> 1; str > 0; str >>= 1) {
      for (int i = threadIdx.x; i < kpow2; i += blockDim.x) {
        int j = i ^ str;
        if (j > i) {
          bool up = ((i & size) == 0);
          unsigned long long ka = keys[i], kb = keys[j];
          unsigned int ia = ids[i], ib = ids[j];
          bool gt = ka > kb || (ka == kb && ia > ib);
          if (gt == up) { keys[i] = kb; keys[j] = ka; ids[i] = ib; ids[j] = ia; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < k; i += blockDim.x) idx_out[i] = (int64_t)ids[i];
}


// ---------------------------------------------------------------------------------------------------
// single-launch radix select
// ---------------------------------------------------------------------------------------------------
constexpr int kSelPasses = 9;
constexpr int kSelMaxBins = 2048;
constexpr int kSelUnroll = 8;

struct FusedState {
  unsigned int hist[kSelPasses][kSelMaxBins];
  unsigned int barrier;        // monotonically increasing arrival counter of the grid barrier
  unsigned int out_count;
  unsigned int ticket;
};

// digit geometry of pass p: passes 0..5 walk the 64-bit key from the top (11,11,11,11,11,9 bits), 6..8 the index (11,11,10)
__device__ __forceinline__ void sel_digit(int pass, int& shift, int& bits, bool& on_idx) {
  on_idx = pass >= 6;
  if (!on_idx) { bits = pass < 5 ? 11 : 9; shift = pass < 5 ? 53 - 11 * pass : 0; }
  else { const int q = pass - 6; bits = q < 2 ? 11 : 10; shift = q < 2 ? 21 - 11 * q : 0; }
}

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    const long long t0 = clock64();
    while (*reinterpret_cast<volatile unsigned int*>(counter) < target) {
      if (clock64() - t0 > 4000000000LL) { printf("sdg select: grid barrier timed out\n"); __trap(); }
    }
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256, 4)
select_fused_kernel(const double* __restrict__ score, int64_t n, int largest, int k, int kpow2, FusedState* st,
                    unsigned long long* __restrict__ sel_key, unsigned int* __restrict__ sel_idx,
                    int64_t* __restrict__ idx_out) {
  extern __shared__ unsigned char smem_raw[];
  unsigned int* sh = reinterpret_cast<unsigned int*>(smem_raw);          // [kSelMaxBins] histogram, later the sort buffers
  __shared__ unsigned int s_warp[8];
  __shared__ unsigned int s_res[3];                                       // digit, candidates above it, candidates in it
  __shared__ bool s_last;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t w0 = (int64_t)blockIdx.x * blockDim.x + (tid & ~31);     // first sample of this warp's 32-wide slice
  unsigned long long pk = 0;
  unsigned int pi = 0, k_rem = (unsigned int)k, target = 0;
  bool done = false;
  for (int pass = 0; pass < kSelPasses && !done; ++pass) {
    int shift, bits;
    bool on_idx;
    sel_digit(pass, shift, bits, on_idx);
    const int nb = 1 << bits;
    const unsigned int mask = (unsigned int)nb - 1u;
    const int hs = shift + bits;                             // first bit above this digit (of the key, or of the index)
    const unsigned long long himask = on_idx ? (hs >= 32 ? 0ULL : (unsigned long long)(~0u << hs))
                                             : (hs >= 64 ? 0ULL : (~0ULL << hs));
    for (int b = tid; b < nb; b += 256) sh[b] = 0;
    __syncthreads();
    // warp-uniform sweep, kSelUnroll independent 8-byte loads in flight per thread; lanes of a warp that hit the same bin
    // (nearly all of them in the first pass: the top 11 bits of a double are its sign and exponent) are combined with
    // match.any so the shared-memory atomic sees one add per distinct bin instead of up to 32 serialised ones
    // (n < 2^32 is an entry-point requirement: sample indices are 32-bit, a wrapped sum is caught by the 64-bit bound test)
    for (int64_t b0 = w0; b0 < n; b0 += (int64_t)kSelUnroll * stride) {
      double v[kSelUnroll];
#pragma unroll
      for (int u = 0; u < kSelUnroll; ++u) {
        const int64_t i = b0 + (int64_t)u * stride + lane;
        v[u] = i < n ? __ldg(score + i) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < kSelUnroll; ++u) {
        const int64_t wb = b0 + (int64_t)u * stride;
        if (wb >= n) break;                                   // warp-uniform
        const int64_t i = wb + lane;
        unsigned long long key; unsigned int id;
        composite(v[u], (unsigned int)i, largest, key, id);
        bool match; unsigned int digit;
        if (!on_idx) {                                        // (pass-uniform branch)
          match = ((key ^ pk) & himask) == 0ULL;             // bits above this digit are already fixed
          digit = (unsigned int)(key >> shift) & mask;
        } else {
          match = key == pk && ((id ^ pi) & (unsigned int)himask) == 0u;
          digit = (id >> shift) & mask;
        }
        match = match && i < n;
        // after the first pass nearly every warp has no candidate left: one vote skips the match / atomic sequence
        if (!__any_sync(0xffffffffu, match)) continue;
        const unsigned int peers = __match_any_sync(0xffffffffu, match ? digit : 0xffffffffu);
        if (match && lane == __ffs(peers) - 1) atomicAdd(&sh[digit], (unsigned int)__popc(peers));
      }
    }
    __syncthreads();
    for (int b = tid; b < nb; b += 256)
      if (sh[b]) atomicAdd(&st->hist[pass][b], sh[b]);
    grid_barrier(&st->barrier, target);
    // every CTA finds the digit of the k_rem-th largest candidate: thread t owns `per` bins, walking DOWN from the top
    const int per = nb >> 8;
    unsigned int mine[8], c = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mine[j] = 0;
      if (j < per) { mine[j] = __ldcg(&st->hist[pass][nb - 1 - (tid * per + j)]); c += mine[j]; }
    }
    unsigned int incl = c;                                 // inclusive scan over threads (thread 0 = the top bins)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    unsigned int base = 0;
    for (int w2 = 0; w2 < wid; ++w2) base += s_warp[w2];
    const unsigned int excl = base + incl - c;
    if (excl < k_rem && k_rem <= excl + c) {               // exactly one thread: the crossing happens inside its bins
      unsigned int above = excl;
      for (int j = 0; j < per; ++j) {
        if (above + mine[j] >= k_rem) { s_res[0] = (unsigned int)(nb - 1 - (tid * per + j)); s_res[1] = above; s_res[2] = mine[j]; break; }
        above += mine[j];
      }
    }
    __syncthreads();
    const unsigned int d = s_res[0];
    k_rem -= s_res[1];
    if (!on_idx) pk |= (unsigned long long)d << shift; else pi |= d << shift;
    done = s_res[2] == k_rem;                              // the whole bin is wanted: no need to split it further
    __syncthreads();
  }
  // compaction: every candidate at or above the threshold composite (pk, pi) -- exactly k of them
  for (int64_t b0 = w0; b0 < n; b0 += (int64_t)kSelUnroll * stride) {
    double v[kSelUnroll];
#pragma unroll
    for (int u = 0; u < kSelUnroll; ++u) {
      const int64_t i = b0 + (int64_t)u * stride + lane;
      v[u] = i < n ? __ldg(score + i) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < kSelUnroll; ++u) {
      const int64_t i = b0 + (int64_t)u * stride + lane;
      if (i >= n) continue;
      unsigned long long key; unsigned int id;
      composite(v[u], (unsigned int)i, largest, key, id);
      if (key > pk || (key == pk && id >= pi)) {
        const unsigned int slot = atomicAdd(&st->out_count, 1u);
        if (slot < (unsigned int)k) { sel_key[slot] = ordered_key(v[u]); sel_idx[slot] = (unsigned int)i; }
      }
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(&st->ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last CTA: bitonic sort of the k survivors by (key, index) ascending, in shared memory
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
  unsigned int* ids = reinterpret_cast<unsigned int*>(keys + kpow2);
  for (int i = tid; i < kpow2; i += 256) {
    keys[i] = i < k ? __ldcg(sel_key + i) : ~0ULL;
    ids[i] = i < k ? __ldcg(sel_idx + i) : ~0u;
  }
  __syncthreads();
  for (int size = 2; size <= kpow2; size <<= 1) {
    for (int str = size >> 1; str > 0; str >>= 1) {
      for (int i = tid; i < kpow2; i += 256) {
        const int j = i ^ str;
        if (j > i) {
          const bool up = (i & size) == 0;
          const unsigned long long ka = keys[i], kb = keys[j];
          const unsigned int ia = ids[i], ib = ids[j];
          const bool gt = ka > kb || (ka == kb && ia > ib);
          if (gt == up) { keys[i] = kb; keys[j] = ka; ids[i] = ib; ids[j] = ia; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < k; i += 256) idx_out[i] = (int64_t)ids[i];
}

constexpr size_t kSelKeyOff = (sizeof(FusedState) + 255) / 256 * 256;   // state block first (both variants)
constexpr size_t kSelIdxOff = kSelKeyOff + sizeof(unsigned long long) * kTopkMax;
constexpr size_t kSelBytes = kSelIdxOff + sizeof(unsigned int) * kTopkMax;
static_assert(sizeof(SelectState) <= kSelKeyOff && sizeof(FusedState) <= kSelKeyOff, "state block");

}  // namespace sdg

using namespace sdg;

extern "C" size_t sdg_topk_workspace_bytes(int64_t) { return kSelBytes; }

extern "C" int sdg_topk_indices(const double* score, int64_t n, int k, int largest, int64_t* idx_out,
                                void* workspace, size_t workspace_bytes, void* stream) {
  SDG_REQUIRE(score && idx_out && workspace, SDG_E_INVALID, "sdg_topk_indices: null pointer");
  SDG_REQUIRE(workspace_bytes >= kSelBytes, SDG_E_INVALID, "sdg_topk_indices: workspace %zu < %zu",
              workspace_bytes, kSelBytes);
  SDG_REQUIRE(((uintptr_t)workspace % 8) == 0, SDG_E_INVALID, "sdg_topk_indices: workspace not 8-byte aligned");
  SDG_REQUIRE(n >= 0 && n < (1LL << 32), SDG_E_INVALID, "sdg_topk_indices: n=%lld out of range", (long long)n);
  SDG_REQUIRE(k >= 0 && k <= n, SDG_E_INVALID, "sdg_topk_indices: k=%d n=%lld", k, (long long)n);
  SDG_REQUIRE(k <= kTopkMax, SDG_E_UNSUPPORTED, "sdg_topk_indices: k=%d > %d", k, kTopkMax);
  if (k == 0) return 0;
  auto* sel_key = reinterpret_cast<unsigned long long*>((char*)workspace + kSelKeyOff);
  auto* sel_idx = reinterpret_cast<unsigned int*>((char*)workspace + kSelIdxOff);
  int kp = 1;
  while (kp < k) kp <<= 1;
  static const int legacy = getenv("SDG_SELECT_LEGACY") ? atoi(getenv("SDG_SELECT_LEGACY")) : 0;   // A/B: the 15-launch form
  if (!legacy) {
    auto* fs = reinterpret_cast<FusedState*>(workspace);
    SDG_CUDA(cudaMemsetAsync(fs, 0, sizeof(FusedState), (cudaStream_t)stream));
    const size_t smem = std::max<size_t>(sizeof(unsigned int) * kSelMaxBins, (size_t)kp * (sizeof(unsigned long long) + sizeof(unsigned int)));
    int dev = 0, sms = kNumSMs;
    SDG_CUDA(cudaGetDevice(&dev));
    SDG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    static std::atomic<unsigned long long> attr_set{0};
    if (dev >= 64 || !((attr_set.load() >> dev) & 1ULL)) {
      SDG_CUDA(cudaFuncSetAttribute(select_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      if (dev < 64) attr_set.fetch_or(1ULL << dev);
    }
    // at most 4 resident CTAs per SM (<= 64 registers, <= 56 KB of shared memory each; 2 when k needs the big sort
    // buffer): the cooperative launch guarantees co-residency, which the in-kernel grid barrier relies on.  The sweeps
    // are latency / issue bound per warp (ncu: 16 warps per SM issue 45 % of the cycles), so residency pays at large n.
    int64_t grid = cdiv(n, 256 * kSelUnroll);
    int per_sm = smem <= 24 * 1024 ? 4 : 2, fit = 0;
    SDG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, select_fused_kernel, 256, smem));
    SDG_REQUIRE(fit >= 1, SDG_E_DEVICE, "sdg_topk_indices: the selection kernel does not fit an SM");
    if (per_sm > fit) per_sm = fit;
    if (grid > (int64_t)per_sm * sms) grid = (int64_t)per_sm * sms;
    if (grid < 1) grid = 1;
    int largest_i = largest, k_i = k, kp_i = kp;
    void* args[] = {(void*)&score, (void*)&n, (void*)&largest_i, (void*)&k_i, (void*)&kp_i, (void*)&fs, (void*)&sel_key,
                    (void*)&sel_idx, (void*)&idx_out};
    SDG_CUDA(cudaLaunchCooperativeKernel((const void*)select_fused_kernel, dim3((unsigned)grid), dim3(256), args, smem,
                                         (cudaStream_t)stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
  }
  auto* st = reinterpret_cast<SelectState*>(workspace);
  SDG_LAUNCH(select_init_kernel, 1, 256, 0, stream, st, (unsigned int)k);
  int grid = stream_grid(n, 256 * 4, 4);
  for (int pass = 0; pass < 12; ++pass)
    SDG_LAUNCH(select_pass_kernel, grid, 256, 0, stream, score, n, largest, pass, st);
  SDG_LAUNCH(select_compact_kernel, grid, 256, 0, stream, score, n, largest, st, sel_key, sel_idx, k);
  size_t smem = (size_t)kp * (sizeof(unsigned long long) + sizeof(unsigned int));
  int threads = kp / 2 < 32 ? 32 : (kp / 2 > 1024 ? 1024 : kp / 2);
  SDG_LAUNCH(select_sort_kernel, 1, threads, smem, stream, sel_key, sel_idx, k, kp, idx_out);
  return 0;
}
