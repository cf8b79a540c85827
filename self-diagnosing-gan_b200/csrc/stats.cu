// stats.cu -- per-sample running statistics and the LDR score stage (fp64, HBM-streaming kernels).
//
// Compiled with -fmad=false: the reference evaluates these expressions in NumPy float64 with separate
// multiplies and adds (diagan-pkg/diagan/utils/plot.py:243-248); contraction into FMA would change the
// last bit and break the bit-exact comparison against np.mean / np.var / np.std.
//
// Rooflines (DESIGN.md "Kernels"): all kernels here are HBM-bound streams.
//   stats_update        4 B read + 4*8 B read + 4*8 B write = 68 B / sample
//   window_moments<f32> 4*T B read + 8 B per requested output / sample   (T*N re-read served by L1/L2)
//   score_floor_min     16 B read + 8 B write / sample / conf
//   score_clip          8 B read + 8 B write / sample / conf
#include <cstdlib>

#include "common.cuh"

namespace sdg {

// ------------------------------------------------------------------------------------------------
// Welford + last + sum|delta| update, two samples per thread (16-byte loads/stores on the fp64 state)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void welford_step(double x, double count, bool first,
                                             double& mean, double& m2, double& last, double& sad) {
  if (!first) sad = sad + fabs(x - last);
  double d = x - mean;
  mean = mean + d / count;               // a true division, like the oracle's NumPy restatement
  m2 = m2 + d * (x - mean);
  last = x;
}

__global__ void __launch_bounds__(256)
stats_update_kernel(const float* __restrict__ snap, double* __restrict__ mean, double* __restrict__ m2,
                    double* __restrict__ last, double* __restrict__ sad, int64_t n, int64_t t, int vec_ok) {
  const double inv = (double)(t + 1);
  const bool first = (t == 0);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t pairs = vec_ok ? n / 2 : 0;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < pairs; p += stride) {
    float2 x = reinterpret_cast<const float2*>(snap)[p];
    double2 me, q, la, sa;
    if (first) {
      me = make_double2(0.0, 0.0); q = me; la = me; sa = me;
    } else {
      me = reinterpret_cast<double2*>(mean)[p];
      q = reinterpret_cast<double2*>(m2)[p];
      la = reinterpret_cast<double2*>(last)[p];
      sa = reinterpret_cast<double2*>(sad)[p];
    }
    welford_step((double)x.x, inv, first, me.x, q.x, la.x, sa.x);
    welford_step((double)x.y, inv, first, me.y, q.y, la.y, sa.y);
    reinterpret_cast<double2*>(mean)[p] = me;
    reinterpret_cast<double2*>(m2)[p] = q;
    reinterpret_cast<double2*>(last)[p] = la;
    reinterpret_cast<double2*>(sad)[p] = sa;
  }
  // scalar tail (odd n, or unaligned shard base)
  for (int64_t i = pairs * 2 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double me = 0.0, q = 0.0, la = 0.0, sa = 0.0;
    if (!first) { me = mean[i]; q = m2[i]; la = last[i]; sa = sad[i]; }
    welford_step((double)snap[i], inv, first, me, q, la, sa);
    mean[i] = me; m2[i] = q; last[i] = la; sad[i] = sa;
  }
}

// ------------------------------------------------------------------------------------------------
// two-pass window moments in NumPy's axis-0 order: thread i owns sample i, warps read rows coalesced
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
window_moments_kernel(const T* __restrict__ snaps, int64_t Tn, int64_t n, int64_t ld,
                      double* __restrict__ mean_out, double* __restrict__ var_out,
                      double* __restrict__ ldrd_out, double* __restrict__ ldr_out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const T* col = snaps + i;
    double acc = 0.0, sad = 0.0, prev = 0.0;
    int64_t t = 0;
    // unrolled by 4: four independent loads in flight, additions still strictly in row order
    for (; t + 4 <= Tn; t += 4) {
      double x0 = (double)col[(t + 0) * ld], x1 = (double)col[(t + 1) * ld];
      double x2 = (double)col[(t + 2) * ld], x3 = (double)col[(t + 3) * ld];
      if (t > 0) sad = sad + fabs(x0 - prev);
      sad = sad + fabs(x1 - x0);
      sad = sad + fabs(x2 - x1);
      sad = sad + fabs(x3 - x2);
      acc = acc + x0; acc = acc + x1; acc = acc + x2; acc = acc + x3;
      prev = x3;
    }
    for (; t < Tn; ++t) {
      double x = (double)col[t * ld];
      if (t > 0) sad = sad + fabs(x - prev);
      acc = acc + x;
      prev = x;
    }
    const double mean = acc / (double)Tn;
    if (mean_out) mean_out[i] = mean;
    if (ldr_out) ldr_out[i] = prev;
    if (ldrd_out) ldrd_out[i] = sad / (double)(Tn - 1);
    if (var_out) {
      double sq = 0.0;
      t = 0;
      for (; t + 4 <= Tn; t += 4) {   // second pass: this thread's column again (L1/L2 resident)
        double d0 = (double)col[(t + 0) * ld] - mean, d1 = (double)col[(t + 1) * ld] - mean;
        double d2 = (double)col[(t + 2) * ld] - mean, d3 = (double)col[(t + 3) * ld] - mean;
        sq = sq + d0 * d0; sq = sq + d1 * d1; sq = sq + d2 * d2; sq = sq + d3 * d3;
      }
      for (; t < Tn; ++t) {
        double d = (double)col[t * ld] - mean;
        sq = sq + d * d;
      }
      var_out[i] = sq / (double)(Tn - 1);
    }
  }
}

// Staged form for windows whose [T, 128-sample] tile fits shared memory twice (the reference's window is 50 or 51 snapshots,
// train_mimicry_phase2.py:92).  The generic kernel above keeps four loads in flight per thread and re-reads its column for
// the centred squares; the long dependent fp64 addition chains (the row order is fixed by the bit-exactness requirement) then
// sit between a thread's loads.  Here the loads are decoupled from the arithmetic: a CTA copies the NEXT tile with cp.async
// (no registers held, 25-50 KB in flight per CTA) while every thread walks its own column of the current tile twice from
// shared memory, in exactly the generic kernel's order -- bit-identical results, every snapshot byte leaves DRAM once.
constexpr int kWmTile = 128;          // samples per tile = threads per CTA

template <int BYTES>
__device__ __forceinline__ void cp_async_n(void* dst_smem, const void* src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
  else if (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}

// NBUF = 2: the next tile is copied while this one is reduced; NBUF = 1: one buffer per CTA and twice the resident CTAs (the
// copies of some CTAs overlap the arithmetic of the others).
template <typename T, int NBUF>
__global__ void __launch_bounds__(kWmTile)
window_moments_tile_kernel(const T* __restrict__ snaps, int Tn, int64_t n, int64_t ld, int vec16,
                           double* __restrict__ mean_out, double* __restrict__ var_out,
                           double* __restrict__ ldrd_out, double* __restrict__ ldr_out) {
  extern __shared__ __align__(16) unsigned char wm_smem[];
  T* bufs[2] = {reinterpret_cast<T*>(wm_smem), reinterpret_cast<T*>(wm_smem) + (size_t)Tn * kWmTile};
  const int tid = threadIdx.x;
  const int64_t n_tiles = (n + kWmTile - 1) / kWmTile;
  constexpr int PER = 16 / (int)sizeof(T);              // samples per 16-byte chunk
  constexpr int CPR = kWmTile / PER;                    // chunks per tile row
  auto issue = [&](int64_t tile, T* dst) {
    const int64_t i0 = tile * kWmTile;
    if (vec16) {                                        // rows 16-byte aligned and n a multiple of PER: whole chunks only
      for (int c = tid; c < Tn * CPR; c += kWmTile) {
        const int t = c / CPR, q = (c % CPR) * PER;
        if (i0 + q < n) cp_async_n<16>(dst + t * kWmTile + q, snaps + (int64_t)t * ld + i0 + q);
      }
    } else if (i0 + tid < n) {
      for (int t = 0; t < Tn; ++t) cp_async_n<(int)sizeof(T)>(dst + t * kWmTile + tid, snaps + (int64_t)t * ld + i0 + tid);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int cur = 0;
  int64_t tile = blockIdx.x;
  if (NBUF == 2 && tile < n_tiles) issue(tile, bufs[0]);
  for (; tile < n_tiles; tile += gridDim.x) {
    if (NBUF == 2) {
      const int64_t next = tile + gridDim.x;
      if (next < n_tiles) issue(next, bufs[cur ^ 1]);
      else asm volatile("cp.async.commit_group;" ::: "memory");     // empty group: the wait below stays uniform
      asm volatile("cp.async.wait_group 1;" ::: "memory");          // everything but the newest group has landed
    } else {
      issue(tile, bufs[0]);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int64_t i = tile * kWmTile + tid;
    if (i < n) {
      const T* col = bufs[cur] + tid;
      double acc = 0.0, sad = 0.0, prev = 0.0, sq = 0.0;
#pragma unroll 5
      for (int t = 0; t < Tn; ++t) {
        const double x = (double)col[t * kWmTile];
        if (t > 0) sad = sad + fabs(x - prev);
        acc = acc + x;
        prev = x;
      }
      const double mean = acc / (double)Tn;
      if (mean_out) mean_out[i] = mean;
      if (var_out) {
#pragma unroll 5
        for (int t = 0; t < Tn; ++t) {
          const double d = (double)col[t * kWmTile] - mean;
          sq = sq + d * d;
        }
      }
      if (ldr_out) ldr_out[i] = prev;
      if (ldrd_out) ldrd_out[i] = sad / (double)(Tn - 1);
      if (var_out) var_out[i] = sq / (double)(Tn - 1);
    }
    __syncthreads();                                    // the tile is free again before the next issue overwrites it
    if (NBUF == 2) cur ^= 1;
  }
}

// ------------------------------------------------------------------------------------------------
// score phase 1: floor + per-conf global min
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (v < __longlong_as_double((long long)old)) {
    unsigned long long assumed = old;
    old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    if (old == assumed) break;
  }
}

__global__ void fill_double_kernel(double* p, int n, double v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

struct ConfTable { double c[128]; };

__device__ __forceinline__ double floor_score(double m, double v, double c, double floor, double m2_over) {
  if (m2_over > 0.0) v = v / m2_over;
  double s = m + c * sqrt(v);
  return s < floor ? floor : s;            // np.clip(a_min=floor): NaN propagates like NumPy
}

// vec_ok: mean / var / score rows are 16-byte aligned and n is even per row -> two samples per 16-byte access, two such
// pairs per iteration (four independent loads in flight per thread)
__global__ void __launch_bounds__(256, 4)      // <= 64 registers: the fp64 sqrt sequence is latency bound, occupancy pays
score_floor_min_kernel(const double* __restrict__ mean, const double* __restrict__ var, int64_t n,
                       ConfTable conf, double floor, double m2_over, double* __restrict__ score,
                       double* __restrict__ mins, int vec_ok) {
  const int j = blockIdx.y;
  const double c = conf.c[j];
  double* out = score + (int64_t)j * n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double lmin = __longlong_as_double(0x7ff0000000000000LL);   // +inf
  const int64_t pairs = vec_ok ? n / 2 : 0;
  const double2* m2p = reinterpret_cast<const double2*>(mean);
  const double2* v2p = reinterpret_cast<const double2*>(var);
  double2* o2p = reinterpret_cast<double2*>(out);
  int64_t p = tid;
  for (; p + stride < pairs; p += 2 * stride) {
    const double2 ma = __ldcs(m2p + p), va = __ldcs(v2p + p), mb = __ldcs(m2p + p + stride), vb = __ldcs(v2p + p + stride);
    double2 sa, sb;
    sa.x = floor_score(ma.x, va.x, c, floor, m2_over); sa.y = floor_score(ma.y, va.y, c, floor, m2_over);
    sb.x = floor_score(mb.x, vb.x, c, floor, m2_over); sb.y = floor_score(mb.y, vb.y, c, floor, m2_over);
    o2p[p] = sa; o2p[p + stride] = sb;
    lmin = sa.x < lmin ? sa.x : lmin; lmin = sa.y < lmin ? sa.y : lmin;
    lmin = sb.x < lmin ? sb.x : lmin; lmin = sb.y < lmin ? sb.y : lmin;
  }
  for (; p < pairs; p += stride) {
    const double2 ma = __ldcs(m2p + p), va = __ldcs(v2p + p);
    double2 sa;
    sa.x = floor_score(ma.x, va.x, c, floor, m2_over); sa.y = floor_score(ma.y, va.y, c, floor, m2_over);
    o2p[p] = sa;
    lmin = sa.x < lmin ? sa.x : lmin; lmin = sa.y < lmin ? sa.y : lmin;
  }
  for (int64_t i = pairs * 2 + tid; i < n; i += stride) {
    const double s = floor_score(mean[i], var[i], c, floor, m2_over);
    out[i] = s;
    lmin = s < lmin ? s : lmin;
  }
  for (int o = 16; o > 0; o >>= 1) {
    double other = __shfl_xor_sync(0xffffffffu, lmin, o);
    lmin = other < lmin ? other : lmin;
  }
  __shared__ double wmin[8];
  if ((threadIdx.x & 31) == 0) wmin[threadIdx.x >> 5] = lmin;
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = wmin[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = wmin[w] < m ? wmin[w] : m;
    atomic_min_double(&mins[j], m);
  }
}

__device__ __forceinline__ double clip_score(double v, double upper, double eps) {
  v = v > upper ? upper : v;
  if (eps > 0.0) v = v < eps ? eps : v;
  return v;
}

__global__ void __launch_bounds__(256)
score_clip_kernel(double* __restrict__ score, int64_t n, const double* __restrict__ mins, double ratio,
                  double eps, int vec_ok) {
  const int j = blockIdx.y;
  const double upper = mins[j] * ratio;
  double* s = score + (int64_t)j * n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t pairs = vec_ok ? n / 2 : 0;
  double2* s2 = reinterpret_cast<double2*>(s);
  int64_t p = tid;
  for (; p + 3 * stride < pairs; p += 4 * stride) {     // four 16-byte loads in flight per thread
    double2 a = s2[p], b = s2[p + stride], c = s2[p + 2 * stride], d = s2[p + 3 * stride];
    a.x = clip_score(a.x, upper, eps); a.y = clip_score(a.y, upper, eps);
    b.x = clip_score(b.x, upper, eps); b.y = clip_score(b.y, upper, eps);
    c.x = clip_score(c.x, upper, eps); c.y = clip_score(c.y, upper, eps);
    d.x = clip_score(d.x, upper, eps); d.y = clip_score(d.y, upper, eps);
    s2[p] = a; s2[p + stride] = b; s2[p + 2 * stride] = c; s2[p + 3 * stride] = d;
  }
  for (; p < pairs; p += stride) {
    double2 a = s2[p];
    a.x = clip_score(a.x, upper, eps); a.y = clip_score(a.y, upper, eps);
    s2[p] = a;
  }
  for (int64_t i = pairs * 2 + tid; i < n; i += stride) s[i] = clip_score(s[i], upper, eps);
}

// one-collective form of the sharded score: gathered = [world][shard + 1] doubles, slot r = rank r's floor-clipped shard
// (padded to `shard`) followed by its local minimum; out[i] = max(min(v, gmin * ratio), eps) with gmin = min over the slots
__global__ void __launch_bounds__(256)
score_clip_gathered_kernel(const double* __restrict__ gathered, int world, int64_t shard, int64_t n, double ratio, double eps,
                           double* __restrict__ out) {
  __shared__ double s_min;
  if (threadIdx.x == 0) {
    double m = gathered[shard];
    for (int r = 1; r < world; ++r) {
      const double v = gathered[(int64_t)r * (shard + 1) + shard];
      m = v < m ? v : m;
    }
    s_min = m;
  }
  __syncthreads();
  const double upper = s_min * ratio;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t r = i / shard;
    double v = gathered[r * (shard + 1) + (i - r * shard)];
    v = v > upper ? upper : v;
    if (eps > 0.0) v = v < eps ? eps : v;
    out[i] = v;
  }
}

}  // namespace sdg

using namespace sdg;

extern "C" int sdg_stats_update(const float* snapshot, double* mean, double* m2, double* last, double* sad,
                                int64_t n, int64_t t, void* stream) {
  SDG_REQUIRE(n >= 0 && t >= 0, SDG_E_INVALID, "sdg_stats_update: n=%lld t=%lld", (long long)n, (long long)t);
  if (n == 0) return 0;          // an empty shard: torch hands out null pointers for empty tensors
  SDG_REQUIRE(snapshot && mean && m2 && last && sad, SDG_E_INVALID, "sdg_stats_update: null pointer");
  auto al = [](const void* p, size_t a) { return ((uintptr_t)p % a) == 0; };
  int vec_ok = al(snapshot, 8) && al(mean, 16) && al(m2, 16) && al(last, 16) && al(sad, 16);
  int grid = stream_grid(cdiv(n, 2), 256);
  SDG_LAUNCH(stats_update_kernel, grid, 256, 0, stream, snapshot, mean, m2, last, sad, n, t, vec_ok);
  return 0;
}

template <typename T>
static int window_moments(const T* snaps, int64_t Tn, int64_t n, int64_t ld, double* mean, double* var,
                          double* ldrd, double* ldr, void* stream) {
  SDG_REQUIRE(snaps, SDG_E_INVALID, "sdg_window_moments: null snapshots");
  SDG_REQUIRE(Tn >= 1 && n >= 0 && ld >= n, SDG_E_INVALID, "sdg_window_moments: T=%lld n=%lld ld=%lld",
              (long long)Tn, (long long)n, (long long)ld);
  if (n == 0) return 0;
  static const int generic = getenv("SDG_MOMENTS_GENERIC") ? atoi(getenv("SDG_MOMENTS_GENERIC")) : 0;
  // SDG_MOMENTS_GENERIC: 1 = the generic kernel, 2 = double-buffered tiles (A/B runs); default = single-buffered tiles
  const int nbuf = generic == 2 ? 2 : 1;
  const size_t smem = nbuf * (size_t)Tn * kWmTile * sizeof(T);
  if (generic != 1 && smem <= 160 * 1024) {
    int dev = 0;
    SDG_CUDA(cudaGetDevice(&dev));
    static std::atomic<unsigned long long> attr_set{0};
    if (dev >= 64 || !((attr_set.load() >> dev) & 1ULL)) {
      SDG_CUDA(cudaFuncSetAttribute((window_moments_tile_kernel<T, 1>), cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      SDG_CUDA(cudaFuncSetAttribute((window_moments_tile_kernel<T, 2>), cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      if (dev < 64) attr_set.fetch_or(1ULL << dev);
    }
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 12 ? 12 : per_sm);
    const int grid = stream_grid(n, kWmTile, per_sm);
    const int vec16 = ((uintptr_t)snaps % 16) == 0 && (ld * sizeof(T)) % 16 == 0 && n % (16 / sizeof(T)) == 0;
    if (nbuf == 2) {
      SDG_LAUNCH((window_moments_tile_kernel<T, 2>), grid, kWmTile, smem, stream, snaps, (int)Tn, n, ld, vec16, mean, var, ldrd, ldr);
    } else {
      SDG_LAUNCH((window_moments_tile_kernel<T, 1>), grid, kWmTile, smem, stream, snaps, (int)Tn, n, ld, vec16, mean, var, ldrd, ldr);
    }
    return 0;
  }
  int grid = stream_grid(n, 256);
  SDG_LAUNCH(window_moments_kernel<T>, grid, 256, 0, stream, snaps, Tn, n, ld, mean, var, ldrd, ldr);
  return 0;
}

extern "C" int sdg_window_moments_f32(const float* snaps, int64_t T, int64_t n, int64_t ld, double* mean,
                                      double* var, double* ldrd, double* ldr, void* stream) {
  return window_moments<float>(snaps, T, n, ld, mean, var, ldrd, ldr, stream);
}

extern "C" int sdg_window_moments_f64(const double* snaps, int64_t T, int64_t n, int64_t ld, double* mean,
                                      double* var, double* ldrd, double* ldr, void* stream) {
  return window_moments<double>(snaps, T, n, ld, mean, var, ldrd, ldr, stream);
}

extern "C" int sdg_score_floor_min(const double* mean, const double* var, int64_t n, const double* conf_host,
                                   int n_conf, double floor, double var_is_m2_over, double* score, double* mins,
                                   void* stream) {
  SDG_REQUIRE(conf_host && mins && (n == 0 || (mean && var && score)), SDG_E_INVALID, "sdg_score_floor_min: null pointer");
  SDG_REQUIRE(n_conf >= 1 && n_conf <= 128, SDG_E_INVALID, "sdg_score_floor_min: n_conf=%d (1..128)", n_conf);
  SDG_REQUIRE(n >= 0, SDG_E_INVALID, "sdg_score_floor_min: n=%lld", (long long)n);
  ConfTable tab;
  for (int j = 0; j < n_conf; ++j) tab.c[j] = conf_host[j];
  SDG_LAUNCH(fill_double_kernel, 1, 128, 0, stream, mins, n_conf, __builtin_inf());
  if (n == 0) return 0;
  // 16-byte accesses need every row of the [n_conf, n] score block aligned: even n (or one row) and aligned bases
  const int vec_ok = ((uintptr_t)mean % 16) == 0 && ((uintptr_t)var % 16) == 0 && ((uintptr_t)score % 16) == 0 &&
                     (n % 2 == 0 || n_conf == 1);
  int gx = stream_grid(cdiv(n, 4), 256, n_conf >= 8 ? 1 : 8);
  SDG_LAUNCH(score_floor_min_kernel, dim3(gx, n_conf), 256, 0, stream, mean, var, n, tab, floor,
             var_is_m2_over, score, mins, vec_ok);
  return 0;
}

extern "C" int sdg_score_clip(double* score, int64_t n, int n_conf, const double* mins, double ratio, double eps,
                              void* stream) {
  SDG_REQUIRE(n_conf >= 1 && n >= 0, SDG_E_INVALID, "sdg_score_clip: n_conf=%d n=%lld", n_conf, (long long)n);
  if (n == 0) return 0;
  SDG_REQUIRE(score && mins, SDG_E_INVALID, "sdg_score_clip: null pointer");
  const int vec_ok = ((uintptr_t)score % 16) == 0 && (n % 2 == 0 || n_conf == 1);
  int gx = stream_grid(cdiv(n, 8), 256, n_conf >= 8 ? 1 : 8);
  SDG_LAUNCH(score_clip_kernel, dim3(gx, n_conf), 256, 0, stream, score, n, mins, ratio, eps, vec_ok);
  return 0;
}

extern "C" int sdg_score_clip_gathered(const double* gathered, int world, int64_t shard_size, int64_t n, double ratio,
                                       double eps, double* out, void* stream) {
  SDG_REQUIRE(gathered && out, SDG_E_INVALID, "sdg_score_clip_gathered: null pointer");
  SDG_REQUIRE(world >= 1 && shard_size >= 1 && n >= 0 && n <= (int64_t)world * shard_size, SDG_E_INVALID,
              "sdg_score_clip_gathered: world=%d shard=%lld n=%lld", world, (long long)shard_size, (long long)n);
  if (n == 0) return 0;
  SDG_LAUNCH(score_clip_gathered_kernel, stream_grid(n, 256, 8), 256, 0, stream, gathered, world, shard_size, n, ratio, eps, out);
  return 0;
}
