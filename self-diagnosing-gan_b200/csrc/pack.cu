// pack.cu -- spectral-norm sigma (eval semantics) and weight pre-packing, run once per recording pass.
//
// Replaces torch-mimicry SpectralNorm.sn_weights as evaluated in eval mode (spectral_norm.py, SURVEY
// 8(a) row a5): W2d = W.view(Cout,-1); v = normalize(u W2d); u' = normalize(v W2d^T);
// sigma = u' W2d v^T; weight used = W / sigma; buffers are NOT updated, so the value is the same for
// all ceil(N/64) batches of a pass (trainer.py:145-154) and is computed exactly once here.
// All layers of a discriminator are processed by the same four launches (layer = blockIdx.y / .x).
#include "kernels.cuh"

namespace sdg {

constexpr float kSnEps = 1e-12f;

constexpr int kSnSlices = 8;   // row slices of u.W (partials summed in a fixed order: sigma stays deterministic)

// vp[z][k] = sum over the rows o of slice z of u[o] * W[o][k]     (thread per column k: coalesced across k)
__global__ void __launch_bounds__(256) sn_uW_kernel(const SnLayer* __restrict__ layers) {
  const SnLayer L = layers[blockIdx.y];
  const int z = blockIdx.z;
  const int per = (L.cout + kSnSlices - 1) / kSnSlices;
  const int o0 = z * per, o1 = min(L.cout, o0 + per);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < L.K; k += gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int o = o0; o < o1; ++o) acc = fmaf(L.u[o], L.W[(int64_t)o * L.K + k], acc);
    L.vp[(int64_t)z * L.K + k] = acc;
  }
}

__device__ float block_sum(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += red[k];
  __syncthreads();
  return t;
}

// v = v_raw / max(||v_raw||, eps)      (one CTA per layer)
__global__ void __launch_bounds__(256) sn_normalize_v_kernel(const SnLayer* __restrict__ layers) {
  __shared__ float red[8];
  const SnLayer L = layers[blockIdx.x];
  float s = 0.f;
  for (int k = threadIdx.x; k < L.K; k += blockDim.x) {
    float a = 0.f;
    for (int z = 0; z < kSnSlices; ++z) a += L.vp[(int64_t)z * L.K + k];      // fixed order
    L.v[k] = a;
    s = fmaf(a, a, s);
  }
  float nrm = sqrtf(block_sum(s, red));
  float d = fmaxf(nrm, kSnEps);
  for (int k = threadIdx.x; k < L.K; k += blockDim.x) L.v[k] = L.v[k] / d;
}

// t[o] = sum_k W[o][k] * v[k]          (one warp per row)
__global__ void __launch_bounds__(256) sn_Wv_kernel(const SnLayer* __restrict__ layers) {
  const SnLayer L = layers[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int o = blockIdx.x * wpb + (threadIdx.x >> 5); o < L.cout; o += gridDim.x * wpb) {
    const float* row = L.W + (int64_t)o * L.K;
    float acc = 0.f;
    for (int k = lane; k < L.K; k += 32) acc = fmaf(row[k], L.v[k], acc);
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) L.t[o] = acc;
  }
}

// u' = t / max(||t||, eps); sigma = u' . t     (one CTA per layer)
__global__ void __launch_bounds__(256) sn_sigma_kernel(const SnLayer* __restrict__ layers, float* __restrict__ sigma) {
  __shared__ float red[8];
  const SnLayer L = layers[blockIdx.x];
  float s = 0.f;
  for (int o = threadIdx.x; o < L.cout; o += blockDim.x) s = fmaf(L.t[o], L.t[o], s);
  float nrm = sqrtf(block_sum(s, red));
  float d = fmaxf(nrm, kSnEps);
  float dot = 0.f;
  for (int o = threadIdx.x; o < L.cout; o += blockDim.x) dot = fmaf(L.t[o] / d, L.t[o], dot);
  dot = block_sum(dot, red);
  if (threadIdx.x == 0) sigma[blockIdx.x] = dot;
}

int sn_sigmas(const SnLayer* layers_dev, const SnLayer* layers_host, int n_layers, float* sigma_dev,
              cudaStream_t s) {
  int maxK = 1, maxC = 1;
  for (int i = 0; i < n_layers; ++i) {
    maxK = layers_host[i].K > maxK ? layers_host[i].K : maxK;
    maxC = layers_host[i].cout > maxC ? layers_host[i].cout : maxC;
  }
  SDG_LAUNCH(sn_uW_kernel, dim3((unsigned)cdiv(maxK, 256), n_layers, kSnSlices), 256, 0, s, layers_dev);
  SDG_LAUNCH(sn_normalize_v_kernel, n_layers, 256, 0, s, layers_dev);
  SDG_LAUNCH(sn_Wv_kernel, dim3((unsigned)cdiv(maxC, 8), n_layers), 256, 0, s, layers_dev);
  SDG_LAUNCH(sn_sigma_kernel, n_layers, 256, 0, s, layers_dev, sigma_dev);
  return 0;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_conv_fp32_kernel(const float* __restrict__ W, const float* __restrict__ sigma, const float* __restrict__ scale,
                      float* __restrict__ wp, int Cout, int Cin, int taps, float mul) {
  const int total = Cout * Cin * taps;
  const float sg = sigma ? sigma[0] : 1.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int o = i % Cout;
    int k = i / Cout;
    int tap = k / Cin, c = k - tap * Cin;
    float w = W[((int64_t)o * Cin + c) * taps + tap];
    if (sigma) w = w / sg;                       // self.weight / sigma
    if (scale) w = w * scale[o];
    if (mul != 1.0f) w = w * mul;                // EqualConv2d: weight * (1/sqrt(Cin k^2))
    wp[i] = w;
  }
}

int pack_conv_fp32(const float* W, const float* sigma, const float* scale, float* wp, int Cout, int Cin, int ks,
                   cudaStream_t s, float mul) {
  int total = Cout * Cin * ks * ks;
  SDG_LAUNCH(pack_conv_fp32_kernel, stream_grid(total, 256), 256, 0, s, W, sigma, scale, wp, Cout, Cin, ks * ks, mul);
  return 0;
}

// element i of a packed [Cout][Kpad] conv weight (the body of pack_conv_h16_kernel and of job type 0 of pack_jobs_kernel)
template <bool F16>
__device__ __forceinline__ void pack_conv_h16_elem(const float* __restrict__ W, bool has_sigma, float sg, const float* __restrict__ scale,
                                                   h16* __restrict__ wb, int Cin, int Kpad, int taps, int ld, int col0, float mul,
                                                   int cin_w, int* ovf, int i) {
  int k = i % Kpad;
  int o = i / Kpad;
  float w = 0.f;
  if (k < taps * Cin) {
    int tap = k / Cin, c = k - tap * Cin;
    if (c < cin_w) {                             // cin_w < Cin: the packed layout pads the input channels with zeros
      w = W[((int64_t)o * cin_w + c) * taps + tap];
      if (has_sigma) w = w / sg;
      if (scale) w = w * scale[o];
      if (mul != 1.0f) w = w * mul;
    }
  }
  if (F16 && !(fabsf(w) <= kF16Max)) range_flag_set(ovf, SDG_RANGE_WEIGHT);
  wb[(int64_t)o * ld + col0 + k] = (h16)(pack_h2<F16>(w, 0.f) & 0xffffu);
}

template <bool F16>
__global__ void __launch_bounds__(256)
pack_conv_h16_kernel(const float* __restrict__ W, const float* __restrict__ sigma, const float* __restrict__ scale,
                     h16* __restrict__ wb, int Cout, int Cin, int Kpad, int taps, int ld, int col0, float mul, int cin_w,
                     int* ovf) {
  const int total = Cout * Kpad;
  const float sg = sigma ? sigma[0] : 1.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
    pack_conv_h16_elem<F16>(W, sigma != nullptr, sg, scale, wb, Cin, Kpad, taps, ld, col0, mul, cin_w, ovf, i);
}

int pack_conv_h16(const float* W, const float* sigma, const float* scale, h16* wb, int Cout, int Cin, int Kpad,
                  int ks, int f16, int ld, int col0, cudaStream_t s, float mul, int cin_w) {
  int total = Cout * Kpad;
  if (cin_w <= 0) cin_w = Cin;
  if (f16) {
    SDG_LAUNCH(pack_conv_h16_kernel<true>, stream_grid(total, 256), 256, 0, s, W, sigma, scale, wb, Cout, Cin, Kpad,
               ks * ks, ld, col0, mul, cin_w, t_range_flag);
  } else {
    SDG_LAUNCH(pack_conv_h16_kernel<false>, stream_grid(total, 256), 256, 0, s, W, sigma, scale, wb, Cout, Cin, Kpad,
               ks * ks, ld, col0, mul, cin_w, t_range_flag);
  }
  return 0;
}

template <bool F16>
__device__ __forceinline__ void pack_pool4_h16_elem(const float* __restrict__ W, float sg, h16* __restrict__ wb, int Cin, int ld,
                                                    int* ovf, int i) {
  const int c = i % Cin;
  const int r = i / Cin;
  const int t = r & 15, o = r >> 4;
  const int a = t >> 2, b = t & 3;
  float acc = 0.f;
  for (int ky = a - 1; ky <= a; ++ky) {
    if (ky < 0 || ky > 2) continue;
    for (int kx = b - 1; kx <= b; ++kx) {
      if (kx < 0 || kx > 2) continue;
      acc += W[((int64_t)o * Cin + c) * 9 + ky * 3 + kx] / sg;         // (W / sigma) as the reference forms it
    }
  }
  if (F16 && !(fabsf(0.25f * acc) <= kF16Max)) range_flag_set(ovf, SDG_RANGE_WEIGHT);
  wb[(int64_t)o * ld + t * Cin + c] = (h16)(pack_h2<F16>(0.25f * acc, 0.f) & 0xffffu);
}

template <bool F16>
__global__ void __launch_bounds__(256)
pack_pool4_h16_kernel(const float* __restrict__ W, const float* __restrict__ sigma, h16* __restrict__ wb, int Cout, int Cin,
                      int ld, int* ovf) {
  const int total = Cout * 16 * Cin;
  const float sg = sigma[0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
    pack_pool4_h16_elem<F16>(W, sg, wb, Cin, ld, ovf, i);
}

int pack_pool4_h16(const float* W, const float* sigma, h16* wb, int Cout, int Cin, int f16, int ld, cudaStream_t s) {
  int total = Cout * 16 * Cin;
  if (f16) { SDG_LAUNCH(pack_pool4_h16_kernel<true>, stream_grid(total, 256), 256, 0, s, W, sigma, wb, Cout, Cin, ld, t_range_flag); }
  else { SDG_LAUNCH(pack_pool4_h16_kernel<false>, stream_grid(total, 256), 256, 0, s, W, sigma, wb, Cout, Cin, ld, t_range_flag); }
  return 0;
}

template <bool F16>
__device__ __forceinline__ void pack_pool4_sc_h16_elem(const float* __restrict__ Wsc, float sg, h16* __restrict__ wb, int Csc,
                                                       int sc_pad, int ld, int col0, int* ovf, int i) {
  const int c = i % sc_pad;
  const int r = i / sc_pad;
  const int t = r & 3, o = r >> 2;
  const float w = c < Csc ? 0.25f * (Wsc[(int64_t)o * Csc + c] / sg) : 0.f;
  if (F16 && !(fabsf(w) <= kF16Max)) range_flag_set(ovf, SDG_RANGE_WEIGHT);
  wb[(int64_t)o * ld + col0 + t * sc_pad + c] = (h16)(pack_h2<F16>(w, 0.f) & 0xffffu);
}

template <bool F16>
__global__ void __launch_bounds__(256)
pack_pool4_sc_h16_kernel(const float* __restrict__ Wsc, const float* __restrict__ sigma, h16* __restrict__ wb, int Cout,
                         int Csc, int sc_pad, int ld, int col0, int* ovf) {
  const int total = Cout * 4 * sc_pad;
  const float sg = sigma[0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
    pack_pool4_sc_h16_elem<F16>(Wsc, sg, wb, Csc, sc_pad, ld, col0, ovf, i);
}

int pack_pool4_sc_h16(const float* Wsc, const float* sigma, h16* wb, int Cout, int Csc, int sc_pad, int f16, int ld, int col0,
                      cudaStream_t s) {
  int total = Cout * 4 * sc_pad;
  if (f16) { SDG_LAUNCH(pack_pool4_sc_h16_kernel<true>, stream_grid(total, 256), 256, 0, s, Wsc, sigma, wb, Cout, Csc, sc_pad, ld, col0, t_range_flag); }
  else { SDG_LAUNCH(pack_pool4_sc_h16_kernel<false>, stream_grid(total, 256), 256, 0, s, Wsc, sigma, wb, Cout, Csc, sc_pad, ld, col0, t_range_flag); }
  return 0;
}

__global__ void scale_vec_kernel(const float* __restrict__ in, const float* __restrict__ sigma, float* __restrict__ out,
                                 int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = sigma ? in[i] / sigma[0] : in[i];
}

__global__ void add_vec_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}

// The 16-bit weight packs of one SNGAN load (W / sigma of every conv, 11 launches for SNGAN-32 before) as ONE launch:
// blockIdx.y = job, grid-stride over its elements in x; the element bodies are the standalone kernels' own.
struct PackJobTable { int n; int pad; PackJob j[kMaxPackJobs]; };
template <bool F16>
__global__ void __launch_bounds__(256) pack_jobs_kernel(const __grid_constant__ PackJobTable tab, int* ovf) {
  const PackJob& J = tab.j[blockIdx.y];
  const float sg = J.sigma[0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < J.total; i += gridDim.x * blockDim.x) {
    if (J.type == PACK_CONV) pack_conv_h16_elem<F16>(J.W, true, sg, nullptr, J.wb, J.Cin, J.Kpad, J.taps, J.ld, J.col0, 1.0f, J.Cin, ovf, i);
    else if (J.type == PACK_POOL4) pack_pool4_h16_elem<F16>(J.W, sg, J.wb, J.Cin, J.ld, ovf, i);
    else pack_pool4_sc_h16_elem<F16>(J.W, sg, J.wb, J.Cin, J.Kpad, J.ld, J.col0, ovf, i);
  }
}

int pack_jobs(const PackJob* jobs, int n, int f16, cudaStream_t s) {
  for (int o = 0; o < n; o += kMaxPackJobs) {
    PackJobTable tab;
    tab.n = n - o < kMaxPackJobs ? n - o : kMaxPackJobs; tab.pad = 0;
    int most = 1;
    for (int i = 0; i < tab.n; ++i) {
      tab.j[i] = jobs[o + i];
      SDG_REQUIRE(tab.j[i].W && tab.j[i].sigma && tab.j[i].wb && tab.j[i].total >= 0, SDG_E_INVALID, "pack_jobs: job %d has a null pointer", o + i);
      if (tab.j[i].total > most) most = tab.j[i].total;
    }
    dim3 grid(stream_grid(most, 256), (unsigned)tab.n);
    if (grid.x > 64) grid.x = 64;                       // x blocks per job: the largest (262 144 elements) takes 16 elements per thread
    if (f16) { SDG_LAUNCH(pack_jobs_kernel<true>, grid, 256, 0, s, tab, t_range_flag); }
    else { SDG_LAUNCH(pack_jobs_kernel<false>, grid, 256, 0, s, tab, t_range_flag); }
  }
  return 0;
}

// out = (a [+ b]) [/ sigma[0]] for up to kMaxVecJobs small vectors in ONE launch (block = job): the bias copies, bias sums and
// W / sigma scalings of a weight load, 18 separate async operations per recording pass before -- which is what a 6 250-sample
// shard's step (1.8 ms on 8 GPUs) notices.  The table travels as a kernel parameter.
struct VecJobTable { int n; int pad; VecJob j[kMaxVecJobs]; };
__global__ void __launch_bounds__(128) vec_jobs_kernel(const __grid_constant__ VecJobTable tab) {
  const VecJob& J = tab.j[blockIdx.x];
  for (int i = threadIdx.x; i < J.n; i += blockDim.x) {
    float v = J.a[i];
    if (J.b) v = v + J.b[i];
    if (J.sigma) v = v / J.sigma[0];
    J.out[i] = v;
  }
}

int vec_jobs(const VecJob* jobs, int n, cudaStream_t s) {
  for (int o = 0; o < n; o += kMaxVecJobs) {
    VecJobTable tab;
    tab.n = n - o < kMaxVecJobs ? n - o : kMaxVecJobs; tab.pad = 0;
    for (int i = 0; i < tab.n; ++i) {
      tab.j[i] = jobs[o + i];
      SDG_REQUIRE(tab.j[i].out && tab.j[i].a && tab.j[i].n >= 0, SDG_E_INVALID, "vec_jobs: job %d has a null pointer", o + i);
    }
    SDG_LAUNCH(vec_jobs_kernel, (unsigned)tab.n, 128, 0, s, tab);
  }
  return 0;
}

int add_vec(const float* a, const float* b, float* out, int n, cudaStream_t s) {
  SDG_LAUNCH(add_vec_kernel, (unsigned)cdiv(n, 256), 256, 0, s, a, b, out, n);
  return 0;
}

int scale_vec(const float* in, const float* sigma, float* out, int n, cudaStream_t s) {
  SDG_LAUNCH(scale_vec_kernel, (unsigned)cdiv(n, 256), 256, 0, s, in, sigma, out, n);
  return 0;
}

__global__ void bn_fold_kernel(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                               float* scale, float* shift, int C, float mul) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) {
    float sc = gamma[i] / sqrtf(var[i] + eps);
    float sh = beta[i] - mean[i] * sc;
    if (mul != 1.0f) { sc *= mul; sh *= mul; }
    scale[i] = sc;
    shift[i] = sh;
  }
}

int bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps, float* scale,
            float* shift, int C, cudaStream_t s, float mul) {
  SDG_LAUNCH(bn_fold_kernel, (unsigned)cdiv(C, 256), 256, 0, s, gamma, beta, mean, var, eps, scale, shift, C, mul);
  return 0;
}

__global__ void permute_fc_kernel(const float* __restrict__ w, float* __restrict__ out, int C, int HW) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;       // NHWC index p*C + c
  if (i < C * HW) {
    int c = i % C, p = i / C;
    out[i] = w[c * HW + p];
  }
}

int permute_fc(const float* w, float* out, int C, int HW, cudaStream_t s) {
  SDG_LAUNCH(permute_fc_kernel, (unsigned)cdiv((int64_t)C * HW, 256), 256, 0, s, w, out, C, HW);
  return 0;
}

template <bool F16>
__global__ void __launch_bounds__(256)
pack_first_superpix_kernel(const float* __restrict__ W, const float* __restrict__ sigma, h16* __restrict__ wb, int Cout,
                           int* ovf) {
  const float sg = sigma ? sigma[0] : 1.f;
  const int total = 2 * Cout * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int row = i >> 6, k = i & 63;
    const int par = row / Cout, o = row - par * Cout;
    float w = 0.f;
    if (k < 36) {
      const int tap = k / 3, c = k - tap * 3;
      const int ky = tap >> 2, j = tap & 3;
      const int kx = j - par;
      if (kx >= 0 && kx <= 2) w = W[((o * 3 + c) * 3 + ky) * 3 + kx] / sg;
    }
    if (F16 && !(fabsf(w) <= kF16Max)) range_flag_set(ovf, SDG_RANGE_WEIGHT);
    wb[i] = (h16)(pack_h2<F16>(w, 0.f) & 0xffffu);
  }
}

int pack_first_superpix_h16(const float* W, const float* sigma, h16* wb, int Cout, int f16, cudaStream_t s) {
  if (f16) { SDG_LAUNCH(pack_first_superpix_kernel<true>, stream_grid(2 * Cout * 64, 256), 256, 0, s, W, sigma, wb, Cout, t_range_flag); }
  else { SDG_LAUNCH(pack_first_superpix_kernel<false>, stream_grid(2 * Cout * 64, 256), 256, 0, s, W, sigma, wb, Cout, t_range_flag); }
  return 0;
}

__global__ void __launch_bounds__(256)
pack_superpix_kernel(const h16* __restrict__ src, h16* __restrict__ dst, int Cout, int Cin, int ty, int txs, int shift) {
  const int txd = txs + shift;
  const int64_t Kd = (int64_t)ty * txd * Cin, Ks = (int64_t)ty * txs * Cin, total = 2LL * Cout * Kd;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int row = (int)(i / Kd);
    const int k = (int)(i - (int64_t)row * Kd);
    const int par = row / Cout, c = row - par * Cout;
    const int tap = k / Cin, ci = k - tap * Cin;
    const int t_y = tap / txd, j = tap - t_y * txd;
    const int tx = j - par * shift;
    dst[i] = (tx >= 0 && tx < txs) ? src[(int64_t)c * Ks + ((int64_t)t_y * txs + tx) * Cin + ci] : (h16)0;
  }
}

int pack_superpix_h16(const h16* src, h16* dst, int Cout, int Cin, int ty, int txs, int shift, cudaStream_t s) {
  const int64_t total = 2LL * Cout * ty * (txs + shift) * Cin;
  SDG_LAUNCH(pack_superpix_kernel, stream_grid(total, 256), 256, 0, s, src, dst, Cout, Cin, ty, txs, shift);
  return 0;
}

int sn_slices() { return kSnSlices; }

}  // namespace sdg
