// drs.cu -- Discriminator Rejection Sampling acceptance pass, one CTA per candidate batch.
//
// Replaces DRS.init_drs / DRS.sub_rejection_sampler (diagan-pkg/diagan/models/drs.py:31-57 and
// diagan-pkg/diagan/trainer/evaluate.py:45-68): float32 arithmetic throughout, like the NumPy original
// (compiled with -fmad=false so `l - eps`, `F - gamma`, the percentile lerp etc. round like NumPy's).
// Latency-bound (n = 128..256 candidates, 4 B read + 5 B written each): the point of the kernel is to
// keep the accept/compact step on the device instead of the reference's per-image .cpu().numpy() loop.
#include "common.cuh"
#include <math_constants.h>

namespace sdg {

constexpr int kDrsMax = 2048;
constexpr int kDrsThreads = 256;

__device__ float block_max(float v, float* scratch) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float m = scratch[0];
  for (int w = 1; w < kDrsThreads / 32; ++w) m = fmaxf(m, scratch[w]);
  __syncthreads();
  return m;
}

__global__ void __launch_bounds__(kDrsThreads)
drs_update_max_kernel(const float* __restrict__ ldr, int n, float* running_max) {
  __shared__ float scratch[kDrsThreads / 32];
  float v = -CUDART_INF_F;
  for (int i = threadIdx.x; i < n; i += kDrsThreads) v = fmaxf(v, ldr[i]);
  float m = block_max(v, scratch);
  if (threadIdx.x == 0 && *running_max < m) *running_max = m;       // drs.py:34-36
}

__global__ void __launch_bounds__(kDrsThreads)
drs_accept_kernel(const float* __restrict__ ldr, int n, int npow2, float* running_max, float eps, float pct,
                  int use_gamma, float gamma_in, const double* __restrict__ psi, float* __restrict__ p_out,
                  uint8_t* __restrict__ accept_out, int32_t* __restrict__ idx_out, int32_t* __restrict__ count_out) {
  __shared__ float Fv[kDrsMax];        // F in sample order
  __shared__ float Fs[kDrsMax];        // F sorted (padded with +inf)
  __shared__ float scratch[kDrsThreads / 32];
  __shared__ int wsum[kDrsThreads / 32];
  __shared__ float s_gamma;

  float v = -CUDART_INF_F;
  for (int i = threadIdx.x; i < n; i += kDrsThreads) v = fmaxf(v, ldr[i]);
  float m = block_max(v, scratch);
  float M = *running_max;
  if (m > M) M = m;                                               // drs.py:39-41
  __syncthreads();
  if (threadIdx.x == 0) *running_max = M;

  for (int i = threadIdx.x; i < npow2; i += kDrsThreads) {
    float f = CUDART_INF_F;
    if (i < n) {
      float l = ldr[i] - M;                                       // drs.py:43
      f = l - logf(1.0f - expf(l - eps));                         // drs.py:45
      Fv[i] = f;
    }
    Fs[i] = f;
  }
  __syncthreads();

  if (!use_gamma) {
    for (int size = 2; size <= npow2; size <<= 1) {
      for (int str = size >> 1; str > 0; str >>= 1) {
        for (int i = threadIdx.x; i < npow2; i += kDrsThreads) {
          int j = i ^ str;
          if (j > i) {
            bool up = ((i & size) == 0);
            float a = Fs[i], b = Fs[j];
            if ((a > b) == up) { Fs[i] = b; Fs[j] = a; }
          }
        }
        __syncthreads();
      }
    }
    if (threadIdx.x == 0) {
      // np.percentile(F, pct) on float32 data: all-float32 linear interpolation (oracle/drs.py percentile_f32)
      float quant = pct / 100.0f;
      float vidx = (float)(n - 1) * quant;
      int lo = (int)floorf(vidx);
      int hi = lo + 1 < n ? lo + 1 : n - 1;
      float t = vidx - (float)lo;
      float a = Fs[lo], b = Fs[hi];
      float d = b - a;
      s_gamma = (t >= 0.5f) ? (b - d * (1.0f - t)) : (a + d * t);
    }
  } else if (threadIdx.x == 0) {
    s_gamma = gamma_in;
  }
  __syncthreads();
  const float gamma = s_gamma;

  // acceptance + ordered compaction, kDrsThreads candidates per sweep
  int base_count = 0;
  for (int start = 0; start < n; start += kDrsThreads) {
    int i = start + threadIdx.x;
    int flag = 0;
    if (i < n) {
      float f = Fv[i] - gamma;                                    // drs.py:52
      float p = 1.0f / (1.0f + expf(-f));                         // drs.py:4-5,53
      if (p_out) p_out[i] = p;
      if (psi) flag = ((double)p > psi[i]) ? 1 : 0;               // drs.py:56
      if (accept_out) accept_out[i] = (uint8_t)flag;
    }
    unsigned ball = __ballot_sync(0xffffffffu, flag);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int within = __popc(ball & ((1u << lane) - 1u));
    if (lane == 0) wsum[w] = __popc(ball);
    __syncthreads();
    int before = 0, total = 0;
    for (int k = 0; k < kDrsThreads / 32; ++k) {
      if (k < w) before += wsum[k];
      total += wsum[k];
    }
    if (flag && idx_out) idx_out[base_count + before + within] = i;
    base_count += total;
    __syncthreads();
  }
  if (threadIdx.x == 0 && count_out) *count_out = base_count;
}

}  // namespace sdg

using namespace sdg;

extern "C" int sdg_drs_update_max(const float* ldr, int n, float* running_max, void* stream) {
  SDG_REQUIRE(ldr && running_max, SDG_E_INVALID, "sdg_drs_update_max: null pointer");
  SDG_REQUIRE(n >= 1, SDG_E_INVALID, "sdg_drs_update_max: n=%d", n);
  SDG_LAUNCH(drs_update_max_kernel, 1, kDrsThreads, 0, stream, ldr, n, running_max);
  return 0;
}

extern "C" int sdg_drs_accept(const float* ldr, int n, float* running_max, float eps, float percentile,
                              int use_gamma, float gamma, const double* psi, float* p_out, uint8_t* accept_out,
                              int32_t* idx_out, int32_t* count_out, void* stream) {
  SDG_REQUIRE(ldr && running_max, SDG_E_INVALID, "sdg_drs_accept: null pointer");
  SDG_REQUIRE(n >= 1 && n <= kDrsMax, SDG_E_UNSUPPORTED, "sdg_drs_accept: n=%d (1..%d)", n, kDrsMax);
  SDG_REQUIRE(use_gamma || (percentile >= 0.f && percentile <= 100.f), SDG_E_INVALID,
              "sdg_drs_accept: percentile=%f", (double)percentile);
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  SDG_LAUNCH(drs_accept_kernel, 1, kDrsThreads, 0, stream, ldr, n, np2, running_max, eps, percentile, use_gamma,
             gamma, psi, p_out, accept_out, idx_out, count_out);
  return 0;
}
