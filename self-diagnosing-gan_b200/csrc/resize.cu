// resize.cu -- the dataset transform of the recording pass on uint8 images, bit-exact with the reference's PIL pipeline.
//
// Replaces transforms.Resize(img_size) + transforms.CenterCrop(img_size) of diagan-pkg/diagan/datasets/transform.py:3-41
// (applied per item, per pass, in DataLoader workers: predefined.py:29-36, color_mnist.py:90-100) by ONE pass over the raw
// uint8 dataset that leaves it resident in HBM at the network's input size; ToTensor + Normalize are fused into the first
// conv's operand load (conv_first.cu).  Arithmetic = Pillow's ImagingResample for 8-bit channels (libImaging/Resample.c,
// bilinear): separable, horizontal pass first into a uint8 intermediate, 22-bit fixed-point coefficients, round-half-up,
// clip to [0,255]; geometry = torchvision Resize(int) / CenterCrop.  Integer arithmetic, so the result is bit-identical to
// what the reference's loader produces (tests: oracle/resize.py pinned against Pillow itself).
//
// One CTA per image: phase 1 resamples horizontally only the rows and columns the cropped output needs into shared memory,
// phase 2 resamples those vertically and writes the size x size x C output.  HBM-bound byte work: each input byte is read
// once (L1 serves the overlapping filter taps), each output byte written once.
#include <cmath>

#include "kernels.cuh"

namespace sdg {

constexpr int kResBits = 32 - 8 - 2;          // Pillow PRECISION_BITS

struct AxisTab {
  std::vector<int> bounds;                    // [out][2] = (first input index, tap count)
  std::vector<int> kk;                        // [out][ksize] fixed-point coefficients
  int ksize = 1;
};

// Pillow precompute_coeffs + normalize_coeffs_8bpc (bilinear, box = whole axis); identity when in == out (PIL copies)
static AxisTab make_axis_tab(int in_size, int out_size) {
  AxisTab t;
  t.bounds.resize((size_t)out_size * 2);
  if (in_size == out_size) {
    t.ksize = 1;
    t.kk.assign(out_size, 1 << kResBits);
    for (int i = 0; i < out_size; ++i) { t.bounds[2 * i] = i; t.bounds[2 * i + 1] = 1; }
    return t;
  }
  const double scale = (double)in_size / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;
  t.ksize = (int)std::ceil(support) * 2 + 1;
  t.kk.assign((size_t)out_size * t.ksize, 0);
  const double ss = 1.0 / filterscale;
  std::vector<double> w(t.ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      double a = (x + xmin - center + 0.5) * ss;
      if (a < 0.0) a = -a;
      w[x] = a < 1.0 ? 1.0 - a : 0.0;
      ww += w[x];
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) w[x] /= ww;
      const double v = w[x] * (double)(1 << kResBits);
      t.kk[(size_t)xx * t.ksize + x] = w[x] < 0 ? (int)(-0.5 + v) : (int)(0.5 + v);
    }
    t.bounds[2 * xx] = xmin;
    t.bounds[2 * xx + 1] = xmax;
  }
  return t;
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kResBits;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

__global__ void __launch_bounds__(256)
resize_crop_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int H, int W, int C, int size, int left, int top,
                      const int* __restrict__ hb, const int* __restrict__ hk, int hks, const int* __restrict__ vb,
                      const int* __restrict__ vk, int vks, int y_first, int y_count) {
  extern __shared__ uint8_t inter[];          // [y_count][size][C]: horizontally resampled rows
  const int64_t n = blockIdx.x;
  const int row_el = size * C;
  const uint8_t* img = in + n * H * W * (int64_t)C;
  for (int idx = threadIdx.x; idx < y_count * row_el; idx += blockDim.x) {
    const int r = idx / row_el, rem = idx - r * row_el;
    const int ox = rem / C, c = rem - ox * C;
    const int xx = left + ox;
    const int xmin = hb[2 * xx], cnt = hb[2 * xx + 1];
    const uint8_t* src = img + ((int64_t)(y_first + r) * W + xmin) * C + c;
    const int* k = hk + xx * hks;
    int acc = 1 << (kResBits - 1);
    for (int x = 0; x < cnt; ++x) acc += (int)__ldg(src + x * C) * k[x];
    inter[idx] = clip8(acc);
  }
  __syncthreads();
  uint8_t* dst = out + n * size * (int64_t)row_el;
  for (int idx = threadIdx.x; idx < size * row_el; idx += blockDim.x) {
    const int oy = idx / row_el, rem = idx - oy * row_el;
    const int yy = top + oy;
    const int ymin = vb[2 * yy] - y_first, cnt = vb[2 * yy + 1];
    const int* k = vk + yy * vks;
    int acc = 1 << (kResBits - 1);
    for (int y = 0; y < cnt; ++y) acc += (int)inter[(ymin + y) * row_el + rem] * k[y];
    dst[idx] = clip8(acc);
  }
}

}  // namespace sdg

using namespace sdg;

extern "C" int sdg_resize_center_crop_u8(const uint8_t* in, int64_t n, int H, int W, int C, int size, uint8_t* out, void* stream) {
  SDG_REQUIRE(in && out, SDG_E_INVALID, "sdg_resize_center_crop_u8: null pointer");
  SDG_REQUIRE(n >= 0 && H >= 1 && W >= 1 && size >= 1 && (C == 1 || C == 3), SDG_E_INVALID,
              "sdg_resize_center_crop_u8: n=%lld H=%d W=%d C=%d size=%d", (long long)n, H, W, C, size);
  if (n == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  // torchvision Resize(int): shorter side -> size, longer -> int(size * long / short); CenterCrop: round((d - size) / 2)
  const int nw = W <= H ? size : (int)((double)size * W / H);
  const int nh = W <= H ? (int)((double)size * H / W) : size;
  SDG_REQUIRE(nw >= size && nh >= size, SDG_E_INVALID, "sdg_resize_center_crop_u8: resized %dx%d smaller than the crop", nw, nh);
  const int left = (int)std::nearbyint((nw - size) / 2.0), top = (int)std::nearbyint((nh - size) / 2.0);
  const AxisTab ht = make_axis_tab(W, nw), vt = make_axis_tab(H, nh);
  const int y_first = vt.bounds[2 * top];
  const int y_last = vt.bounds[2 * (top + size - 1)] + vt.bounds[2 * (top + size - 1) + 1];
  const int y_count = y_last - y_first;
  const size_t smem = (size_t)y_count * size * C;
  SDG_REQUIRE(smem <= 200 * 1024, SDG_E_UNSUPPORTED, "sdg_resize_center_crop_u8: %d rows x %d x %d intermediate exceeds shared "
              "memory", y_count, size, C);
  SDG_REQUIRE(n < (1LL << 31), SDG_E_UNSUPPORTED, "sdg_resize_center_crop_u8: n too large");
  // coefficient tables: one stream-ordered allocation [hb | hk | vb | vk]
  const size_t n_hb = ht.bounds.size(), n_hk = ht.kk.size(), n_vb = vt.bounds.size(), n_vk = vt.kk.size();
  std::vector<int> host(n_hb + n_hk + n_vb + n_vk);
  std::copy(ht.bounds.begin(), ht.bounds.end(), host.begin());
  std::copy(ht.kk.begin(), ht.kk.end(), host.begin() + n_hb);
  std::copy(vt.bounds.begin(), vt.bounds.end(), host.begin() + n_hb + n_hk);
  std::copy(vt.kk.begin(), vt.kk.end(), host.begin() + n_hb + n_hk + n_vb);
  int* tab = nullptr;
  SDG_CUDA(cudaMallocAsync(&tab, host.size() * sizeof(int), s));
  SDG_CUDA(cudaMemcpyAsync(tab, host.data(), host.size() * sizeof(int), cudaMemcpyHostToDevice, s));   // pageable: staged before return
  // function attributes are per device: one bit per device id (a single process may drive several GPUs)
  static std::atomic<unsigned long long> attr_set{0};
  int dev = 0;
  SDG_CUDA(cudaGetDevice(&dev));
  if (dev >= 64 || !((attr_set.load() >> dev) & 1ULL)) {
    SDG_CUDA(cudaFuncSetAttribute(resize_crop_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (dev < 64) attr_set.fetch_or(1ULL << dev);
  }
  SDG_LAUNCH(resize_crop_u8_kernel, (unsigned)n, 256, smem, s, in, out, H, W, C, size, left, top, tab, tab + n_hb, ht.ksize,
             tab + n_hb + n_hk, tab + n_hb + n_hk + n_vb, vt.ksize, y_first, y_count);
  SDG_CUDA(cudaFreeAsync(tab, s));
  return 0;
}
