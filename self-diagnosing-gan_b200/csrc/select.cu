// select.cu -- top-k index selection with the tie-break of a stable argsort.
//
// Replaces np.argsort(w)[-k:] / [:k] (eval_gan_drs_with_index.py:97-99, eval_gan_with_index.py:93-95,
// plot.py:100-101).  A stable ascending argsort orders samples by the composite key (score, index),
// which is a strict total order; the last / first k of it are therefore found by a radix select over
// the 96-bit composite (64-bit order-preserving image of the double, 32-bit index): 12 digit passes of
// 8 bits, each a grid-wide histogram over the still-matching candidates followed by a prefix scan of
// the 256 bins done by the last block to finish (ticket), then one compaction pass and a bitonic sort
// of the k survivors in shared memory.  HBM-bound: 8 B read per sample per pass.
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

constexpr size_t kSelKeyOff = 2048;   // SelectState fits in the first 2 KB
constexpr size_t kSelIdxOff = kSelKeyOff + sizeof(unsigned long long) * kTopkMax;
constexpr size_t kSelBytes = kSelIdxOff + sizeof(unsigned int) * kTopkMax;
static_assert(sizeof(SelectState) <= kSelKeyOff, "state block");

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
  auto* st = reinterpret_cast<SelectState*>(workspace);
  auto* sel_key = reinterpret_cast<unsigned long long*>((char*)workspace + kSelKeyOff);
  auto* sel_idx = reinterpret_cast<unsigned int*>((char*)workspace + kSelIdxOff);
  SDG_LAUNCH(select_init_kernel, 1, 256, 0, stream, st, (unsigned int)k);
  int grid = stream_grid(n, 256 * 4, 4);
  for (int pass = 0; pass < 12; ++pass)
    SDG_LAUNCH(select_pass_kernel, grid, 256, 0, stream, score, n, largest, pass, st);
  SDG_LAUNCH(select_compact_kernel, grid, 256, 0, stream, score, n, largest, st, sel_key, sel_idx, k);
  int kp = 1;
  while (kp < k) kp <<= 1;
  size_t smem = (size_t)kp * (sizeof(unsigned long long) + sizeof(unsigned int));
  int threads = kp / 2 < 32 ? 32 : (kp / 2 > 1024 ? 1024 : kp / 2);
  SDG_LAUNCH(select_sort_kernel, 1, threads, smem, stream, sel_key, sel_idx, k, kp, idx_out);
  return 0;
}
