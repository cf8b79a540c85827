// dnet.cu -- discriminator engines behind the C ABI: context, weight loading, the recording-pass forward.
//
// Replaces `netD(real_data)` inside LogTrainer._get_logit (diagan-pkg/diagan/trainer/trainer.py:145-154)
// and DRS.get_fake_samples_and_ldr (diagan-pkg/diagan/models/drs.py:21-29) for
//   * torch-mimicry SNGANDiscriminator32 / 64 (constructed at predefined_models.py:14,38,50,76,88)
//   * MNIST_DCGAN_Discriminator in eval mode (diagan-pkg/diagan/models/mnist.py:155-223)
// Samples are processed in chunks that bound the activation scratch; within a chunk every layer is one
// launch over all samples of the chunk.
#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cmath>
#include <cstring>

#include "kernels.cuh"

namespace sdg {

static thread_local std::string t_error;
thread_local int* t_range_flag = nullptr;
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  t_error = buf;
}

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need) {
    if (need <= bytes) return 0;
    if (p) { cudaFree(p); p = nullptr; bytes = 0; }
    SDG_CUDA(cudaMalloc(&p, need));
    bytes = need;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
  }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct ConvLayer {
  int cout = 0, cin = 0, ks = 0, stride = 1;
  int kpad = 0;              // bf16 K (taps*cin rounded up to 64)
  DevBuf w32, w16, bias;
  DevBuf w3, bias_sum;       // 16-bit path: fp32 [Cout][3] shortcut weights (DBlockOptimized), conv + shortcut bias
  int ktot = 0;              // 16-bit path: K columns of w16 (conv taps + folded shortcut columns)
  int pool4 = 0;             // 16-bit path: packed for the 4x4 stride-2 form of conv3x3 + avg_pool2d
  // super-pixel form of a Cout = 64 layer (two adjacent output pixels = one 128-channel GEMM pixel):
  // w16s [128][ty * (tx + shift) * Cin], bias2 / w3s = the 64-channel vectors twice
  int superpix = 0;
  DevBuf w16s, bias2, w3s;
  DevBuf w16f;               // SNGAN-32 block1.c2 for the one-kernel block 1 (conv_b1fused.cu): w16's 2048 columns + the shortcut chunk
  bool has_bias = false;
  void release_all() {
    w32.release(); w16.release(); bias.release(); w3.release(); bias_sum.release();
    w16s.release(); bias2.release(); w3s.release(); w16f.release();
  }
};

struct BlockSpec { int kind, cin, cout, down; };   // kind 0 = DBlockOptimized, 1 = DBlock

}  // namespace sdg

using namespace sdg;

struct sdg_ctx {
  int device = 0;
  int arch = 0;
  int precision = SDG_PREC_FP32;
  int inplace_relu = 1;
  int64_t chunk = 0;
  int size = 0;                          // input H = W
  std::vector<BlockSpec> blocks;
  std::vector<ConvLayer> convs;          // forward order (SNGAN: c1,c2,c_sc per block; DCGAN: 6 convs)
  std::vector<int> block_first_conv;     // index of a block's c1 in `convs`
  std::vector<int> block_has_sc;
  DevBuf head_w, head_b;                 // fp32
  int head_len = 0;
  DevBuf sigma;                          // [n_layers]
  int n_layers = 0;
  DevBuf sn_table, sn_scratch, bn_scratch;
  DevBuf buf[7], xin, xpool;             // activation scratch
  bool loaded = false;
  // StyleGAN2 discriminator (fp32): ResBlock channel pairs, reference batch size for minibatch-stddev
  std::vector<std::pair<int, int>> sg2_blocks;
  int sg2_batch = 4;
  DevBuf sg2_sd;
  int* range_flag = nullptr;             // caller-owned device int32 (sdg_ctx_set_range_flag) or null
  // optional timing of the dominant kernel (block1.c2) with events on the launching stream
  bool profile = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  size_t prof_used = 0;
  double prof_flops = 0.0;
};

static int64_t sg2_buf_elems(const sdg_ctx* c);
static int forward_stylegan2_fp32(sdg_ctx* c, const void* x, int layout, int64_t nb, float* logits, cudaStream_t s);
static int forward_stylegan2_h16(sdg_ctx* c, const void* x, int layout, int64_t nb, float* logits, cudaStream_t s);
static int sg2_ensure_h16(sdg_ctx* c, int64_t chunk);
static int64_t sg2_h16_bytes_per_sample(const sdg_ctx* c);

static int prof_begin(sdg_ctx* c, cudaStream_t s) {
  if (!c->profile) return 0;
  if (c->prof_used == c->prof_events.size()) {
    cudaEvent_t a, b;
    SDG_CUDA(cudaEventCreate(&a));
    SDG_CUDA(cudaEventCreate(&b));
    c->prof_events.push_back({a, b});
  }
  SDG_CUDA(cudaEventRecord(c->prof_events[c->prof_used].first, s));
  return 0;
}

static int prof_end(sdg_ctx* c, cudaStream_t s, double flops) {
  if (!c->profile) return 0;
  SDG_CUDA(cudaEventRecord(c->prof_events[c->prof_used].second, s));
  c->prof_used++;
  c->prof_flops += flops;
  return 0;
}

static int64_t max_act_elems(const sdg_ctx* c) {
  // largest per-sample activation tensor (elements)
  int64_t m = 0;
  if (c->arch == SDG_ARCH_DCGAN32) return c->precision == SDG_PREC_FP32 ? 16 * 16 * 32 : 16 * 16 * 64;
  int hw = c->size;
  for (auto& b : c->blocks) {
    int hidden = b.kind == 0 ? b.cout : b.cin;
    int64_t a = (int64_t)hw * hw * (hidden > b.cout ? hidden : b.cout);
    m = a > m ? a : m;
    if (b.down) hw /= 2;
  }
  return m;
}

extern "C" const char* sdg_last_error(void) { return t_error.c_str(); }
extern "C" int sdg_abi_version(void) { return SDG_ABI_VERSION; }
extern "C" int64_t sdg_launch_count(void) { return (int64_t)g_launches.load(); }
extern "C" void sdg_launch_count_reset(void) { g_launches.store(0); }

extern "C" int sdg_ctx_create(int device, sdg_ctx** out) {
  SDG_REQUIRE(out, SDG_E_INVALID, "sdg_ctx_create: null out");
  int count = 0;
  SDG_CUDA(cudaGetDeviceCount(&count));
  SDG_REQUIRE(device >= 0 && device < count, SDG_E_INVALID, "sdg_ctx_create: device %d of %d", device, count);
  SDG_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  SDG_CUDA(cudaGetDeviceProperties(&prop, device));
  SDG_REQUIRE(prop.major == 10, SDG_E_DEVICE, "sdg_ctx_create: device %d is sm_%d%d; this library is sm_100a only",
              device, prop.major, prop.minor);
  sdg_ctx* c = new sdg_ctx();
  c->device = device;
  *out = c;
  return 0;
}

extern "C" int sdg_ctx_destroy(sdg_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  for (auto& l : c->convs) l.release_all();
  c->head_w.release(); c->head_b.release(); c->sigma.release();
  c->sn_table.release(); c->sn_scratch.release(); c->bn_scratch.release();
  for (auto& b : c->buf) b.release();
  c->xin.release(); c->xpool.release(); c->sg2_sd.release();
  delete c;
  return 0;
}

extern "C" int sdg_ctx_set_range_flag(sdg_ctx* c, int32_t* device_flag) {
  SDG_REQUIRE(c, SDG_E_INVALID, "sdg_ctx_set_range_flag: null ctx");
  SDG_REQUIRE(((uintptr_t)device_flag % 4) == 0, SDG_E_INVALID, "sdg_ctx_set_range_flag: misaligned flag");
  c->range_flag = device_flag;
  return 0;
}

extern "C" int sdg_ctx_set_chunk(sdg_ctx* c, int64_t samples_per_chunk) {
  SDG_REQUIRE(c && samples_per_chunk >= 0, SDG_E_INVALID, "sdg_ctx_set_chunk: bad argument");
  c->chunk = samples_per_chunk;
  return 0;
}

static void sngan_spec(int arch, std::vector<BlockSpec>& blocks, int& size, int& ndf) {
  blocks.clear();
  if (arch == SDG_ARCH_SNGAN32) {
    blocks = {{0, 3, 128, 1}, {1, 128, 128, 1}, {1, 128, 128, 0}, {1, 128, 128, 0}};
    size = 32; ndf = 128;
  } else {
    blocks = {{0, 3, 64, 1}, {1, 64, 128, 1}, {1, 128, 256, 1}, {1, 256, 512, 1}, {1, 512, 1024, 1}};
    size = 64; ndf = 1024;
  }
}

extern "C" int sdg_sngan_load(sdg_ctx* c, int arch, int n_layers, const float* const* W, const float* const* b,
                              const float* const* u, int precision, int inplace_relu, void* stream) {
  SDG_REQUIRE(c && W && b && u, SDG_E_INVALID, "sdg_sngan_load: null pointer");
  SDG_REQUIRE(arch == SDG_ARCH_SNGAN32 || arch == SDG_ARCH_SNGAN64, SDG_E_INVALID, "sdg_sngan_load: arch=%d", arch);
  SDG_REQUIRE(precision == SDG_PREC_FP32 || precision == SDG_PREC_BF16 || precision == SDG_PREC_FP16, SDG_E_INVALID,
              "sdg_sngan_load: precision=%d", precision);
  cudaStream_t s = (cudaStream_t)stream;
  SDG_CUDA(cudaSetDevice(c->device));
  RangeScope range_scope(c->range_flag);
  int ndf = 0;
  sngan_spec(arch, c->blocks, c->size, ndf);

  // enumerate layers in forward order
  struct L { int cout, cin, ks; };
  std::vector<L> ls;
  c->block_first_conv.clear(); c->block_has_sc.clear();
  for (auto& bl : c->blocks) {
    int hidden = bl.kind == 0 ? bl.cout : bl.cin;
    c->block_first_conv.push_back((int)ls.size());
    ls.push_back({hidden, bl.cin, 3});
    ls.push_back({bl.cout, hidden, 3});
    int sc = (bl.kind == 0) || bl.cin != bl.cout || bl.down;
    c->block_has_sc.push_back(sc);
    if (sc) ls.push_back({bl.cout, bl.cin, 1});
  }
  const int n_convs = (int)ls.size();
  SDG_REQUIRE(n_layers == n_convs + 1, SDG_E_INVALID, "sdg_sngan_load: arch %d has %d layers, got %d", arch,
              n_convs + 1, n_layers);
  for (int i = 0; i < n_layers; ++i)
    SDG_REQUIRE(W[i] && b[i] && u[i], SDG_E_INVALID, "sdg_sngan_load: layer %d has a null pointer", i);

  c->arch = arch; c->precision = precision; c->inplace_relu = inplace_relu ? 1 : 0; c->n_layers = n_layers;
  c->loaded = false;
  if (precision != SDG_PREC_FP32) { int rc = conv_tc_init(c->device); if (rc) return rc; }

  // ---- sigma for every layer: one batched power iteration ----
  std::vector<SnLayer> tab(n_layers);
  size_t scratch_floats = 0;
  for (int i = 0; i < n_layers; ++i) {
    int cout = i < n_convs ? ls[i].cout : 1;
    int K = i < n_convs ? ls[i].cin * ls[i].ks * ls[i].ks : ndf;
    tab[i].W = W[i]; tab[i].u = u[i]; tab[i].cout = cout; tab[i].K = K;
    scratch_floats += (size_t)K * (1 + sn_slices()) + cout;
  }
  { int rc = c->sn_scratch.ensure(scratch_floats * sizeof(float)); if (rc) return rc; }
  { int rc = c->sn_table.ensure(sizeof(SnLayer) * n_layers); if (rc) return rc; }
  { int rc = c->sigma.ensure(sizeof(float) * n_layers); if (rc) return rc; }
  float* sp = c->sn_scratch.as<float>();
  for (int i = 0; i < n_layers; ++i) {
    tab[i].v = sp; sp += tab[i].K; tab[i].t = sp; sp += tab[i].cout;
    tab[i].vp = sp; sp += (size_t)tab[i].K * sn_slices();
  }
  SDG_CUDA(cudaMemcpyAsync(c->sn_table.p, tab.data(), sizeof(SnLayer) * n_layers, cudaMemcpyHostToDevice, s));
  { int rc = sn_sigmas(c->sn_table.as<SnLayer>(), tab.data(), n_layers, c->sigma.as<float>(), s); if (rc) return rc; }

  // ---- pack W / sigma ----
  if ((int)c->convs.size() != n_convs) {
    for (auto& l : c->convs) l.release_all();
    c->convs.assign(n_convs, ConvLayer());
  }
  const float* sig = c->sigma.as<float>();
  // bias copies, bias sums, fp32 W / sigma vectors: collected here and run as ONE launch (vec_jobs) ahead of the weight packs
  std::vector<VecJob> vj;
  auto vjob = [&](void* out, const float* a, const float* b2, const float* sg, int n) {
    vj.push_back(VecJob{(float*)out, a, b2, sg, n, 0});
  };
  for (int i = 0; i < n_convs; ++i) {
    ConvLayer& l = c->convs[i];
    l.cout = ls[i].cout; l.cin = ls[i].cin; l.ks = ls[i].ks; l.stride = 1; l.has_bias = true;
    int K = l.cin * l.ks * l.ks;
    l.kpad = (K + 63) / 64 * 64;
    { int rc = l.bias.ensure(sizeof(float) * l.cout); if (rc) return rc; }
    vjob(l.bias.p, b[i], nullptr, nullptr, l.cout);
    if (precision == SDG_PREC_FP32) {
      { int rc = l.w32.ensure(sizeof(float) * K * l.cout); if (rc) return rc; }
      int rc = pack_conv_fp32(W[i], sig + i, nullptr, l.w32.as<float>(), l.cout, l.cin, l.ks, s);
      if (rc) return rc;
    }
  }
  c->head_len = ndf;
  { int rc = c->head_w.ensure(sizeof(float) * ndf); if (rc) return rc; }
  { int rc = c->head_b.ensure(sizeof(float)); if (rc) return rc; }
  vjob(c->head_w.p, W[n_convs], nullptr, sig + n_convs, ndf);
  vjob(c->head_b.p, b[n_convs], nullptr, nullptr, 1);
  if (precision != SDG_PREC_FP32) {
    // biases of c2 and c_sc summed; the 1x1 shortcut of DBlockOptimized kept as fp32 [Cout][3] / sigma
    for (size_t bi = 0; bi < c->blocks.size(); ++bi) {
      const int i2 = c->block_first_conv[bi] + 1, isc = i2 + 1;
      ConvLayer& l2 = c->convs[i2];
      { int rc = l2.bias_sum.ensure(sizeof(float) * l2.cout); if (rc) return rc; }
      if (c->block_has_sc[bi]) {
        ConvLayer& lsc = c->convs[isc];
        if (c->blocks[bi].kind != 1) {
          { int rc = lsc.w3.ensure(sizeof(float) * 3 * lsc.cout); if (rc) return rc; }
          vjob(lsc.w3.p, W[isc], nullptr, sig + isc, 3 * lsc.cout);
        }
        vjob(l2.bias_sum.p, b[i2], b[isc], nullptr, l2.cout);
      } else {
        vjob(l2.bias_sum.p, b[i2], nullptr, nullptr, l2.cout);
      }
    }
  }
  { int rc = vec_jobs(vj.data(), (int)vj.size(), s); if (rc) return rc; }
  if (precision != SDG_PREC_FP32) {
    // 16-bit path: c1 as is; c2 with the block's 1x1 shortcut conv folded in as extra K columns (res blocks) or
    // kept as fp32 [Cout][3] for the 3-FMA epilogue (DBlockOptimized).
    const int f16 = precision == SDG_PREC_FP16;
    // the W / sigma packs are collected and run as ONE launch (pack_jobs); a consumer of packed weights flushes the list first
    std::vector<PackJob> pj;
    auto pjob = [&](int type, const float* Wp, const float* sg, h16* wb, int Cin_, int Kpad_, int taps, int ld, int col0, int total) {
      pj.push_back(PackJob{Wp, sg, wb, type, Cin_, Kpad_, taps, ld, col0, total, 0});
    };
    auto pflush = [&]() -> int {
      int rc = pj.empty() ? 0 : pack_jobs(pj.data(), (int)pj.size(), f16, s);
      pj.clear();
      return rc;
    };
    for (size_t bi = 0; bi < c->blocks.size(); ++bi) {
      const int i1 = c->block_first_conv[bi], i2 = i1 + 1, isc = i1 + 2;
      const bool has_sc = c->block_has_sc[bi] != 0;
      ConvLayer& l1 = c->convs[i1];
      ConvLayer& l2 = c->convs[i2];
      l1.ktot = l1.kpad;
      { int rc = l1.w16.ensure(sizeof(h16) * (size_t)l1.ktot * l1.cout); if (rc) return rc; }
      pjob(PACK_CONV, W[i1], sig + i1, l1.w16.as<h16>(), l1.cin, l1.kpad, l1.ks * l1.ks, l1.ktot, 0, l1.cout * l1.kpad);
      const bool fold = has_sc && c->blocks[bi].kind == 1;
      // pooled blocks: conv3x3 + avg_pool2d(2) == a 4x4 stride-2 conv with summed weights (16/36 of the MACs)
      static const int use_pool4 = getenv("SDG_POOL4") ? atoi(getenv("SDG_POOL4")) : 1;
      l2.pool4 = (use_pool4 && c->blocks[bi].down) ? 1 : 0;
      const int sc_cols = fold ? (l2.pool4 ? 4 : 1) * c->convs[isc].kpad : 0;
      const int main_cols = l2.pool4 ? 16 * l2.cin : l2.kpad;
      l2.ktot = main_cols + sc_cols;
      { int rc = l2.w16.ensure(sizeof(h16) * (size_t)l2.ktot * l2.cout); if (rc) return rc; }
      if (l2.pool4) pjob(PACK_POOL4, W[i2], sig + i2, l2.w16.as<h16>(), l2.cin, 0, 9, l2.ktot, 0, l2.cout * 16 * l2.cin);
      else pjob(PACK_CONV, W[i2], sig + i2, l2.w16.as<h16>(), l2.cin, l2.kpad, l2.ks * l2.ks, l2.ktot, 0, l2.cout * l2.kpad);
      if (has_sc) {
        ConvLayer& lsc = c->convs[isc];
        if (fold && l2.pool4)
          pjob(PACK_POOL4_SC, W[isc], sig + isc, l2.w16.as<h16>(), lsc.cin, lsc.kpad, 1, l2.ktot, main_cols, lsc.cout * 4 * lsc.kpad);
        else if (fold)
          pjob(PACK_CONV, W[isc], sig + isc, l2.w16.as<h16>(), lsc.cin, lsc.kpad, 1, l2.ktot, l2.kpad, lsc.cout * lsc.kpad);
      }
      // ---- one-kernel block 1 of SNGAN-32 / SNGAN-64: the same c2 weights with the shortcut / bias chunk appended ----
      if (c->blocks[bi].kind == 0 && l2.pool4 && l2.cout == l2.cin && l1.cin == 3 &&
          ((arch == SDG_ARCH_SNGAN32 && l2.cout == 128) || (arch == SDG_ARCH_SNGAN64 && l2.cout == 64))) {
        const int ch = l2.cout;
        { int rc = l2.w16f.ensure(sizeof(h16) * (size_t)b1_fused_w2_elems(ch)); if (rc) return rc; }
        pjob(PACK_POOL4, W[i2], sig + i2, l2.w16f.as<h16>(), l2.cin, 0, 9, b1_fused_w2_ld(ch), 0, l2.cout * 16 * l2.cin);
        { int rc = b1_fused_pack(nullptr, c->convs[isc].w3.as<float>(), l2.bias_sum.as<float>(), l2.w16f.as<h16>(), ch, f16, s);
          if (rc) return rc; }
      }
      // ---- super-pixel forms of the Cout = 64 layers (SNGAN-64 block1.c2 and block2.c1) ----
      static const int use_superpix = getenv("SDG_SUPERPIX") ? atoi(getenv("SDG_SUPERPIX")) : 1;
      auto dup = [&](DevBuf& dst, const float* src, int n_el) -> int {
        { int rc = dst.ensure(sizeof(float) * 2 * n_el); if (rc) return rc; }
        SDG_CUDA(cudaMemcpyAsync(dst.p, src, sizeof(float) * n_el, cudaMemcpyDeviceToDevice, s));
        SDG_CUDA(cudaMemcpyAsync(dst.as<float>() + n_el, src, sizeof(float) * n_el, cudaMemcpyDeviceToDevice, s));
        return 0;
      };
      l1.superpix = l2.superpix = 0;
      if (use_superpix && c->blocks[bi].kind == 0 && l1.cout == 64 && l1.cin == 3) {
        // block1.c1 from the image bytes: 3 x 3 taps -> 3 x 4 taps, x-stride 2, [128][64] weight tile
        { int rc = l1.w16s.ensure(sizeof(h16) * 128 * 64); if (rc) return rc; }
        { int rc = pack_first_superpix_h16(W[i1], sig + i1, l1.w16s.as<h16>(), 64, f16, s); if (rc) return rc; }
        l1.superpix = 1;
      }
      if (use_superpix && c->blocks[bi].kind == 0 && l2.cout == 64 && l2.cin == 64 && l2.pool4) {
        // block1.c2 in the 4x4 stride-2 form: 4 x 4 taps, x-stride 2 -> 4 x 6 taps, x-stride 4
        { int rc = l2.w16s.ensure(sizeof(h16) * 128 * 4 * 6 * l2.cin); if (rc) return rc; }
        { int rc = pflush(); if (rc) return rc; }        // reads the packed l2.w16
        { int rc = pack_superpix_h16(l2.w16.as<h16>(), l2.w16s.as<h16>(), 64, l2.cin, 4, 4, 2, s); if (rc) return rc; }
        { int rc = dup(l2.bias2, l2.bias_sum.as<float>(), 64); if (rc) return rc; }
        { int rc = dup(c->convs[isc].w3s, c->convs[isc].w3.as<float>(), 64 * 3); if (rc) return rc; }
        l2.superpix = 1;
      }
      if (use_superpix && c->blocks[bi].kind == 1 && l1.cout == 64 && l1.cin == 64) {
        // a plain 3x3 conv: 3 x 3 taps -> 3 x 4 taps, x-stride 2
        { int rc = l1.w16s.ensure(sizeof(h16) * 128 * 3 * 4 * l1.cin); if (rc) return rc; }
        { int rc = pflush(); if (rc) return rc; }        // reads the packed l1.w16
        { int rc = pack_superpix_h16(l1.w16.as<h16>(), l1.w16s.as<h16>(), 64, l1.cin, 3, 3, 1, s); if (rc) return rc; }
        { int rc = dup(l1.bias2, l1.bias.as<float>(), 64); if (rc) return rc; }
        l1.superpix = 1;
      }
    }
    { int rc = pflush(); if (rc) return rc; }
  }
  c->loaded = true;
  return 0;
}

extern "C" int sdg_sngan_sigmas(sdg_ctx* c, float* sigma_out, void* stream) {
  SDG_REQUIRE(c && sigma_out, SDG_E_INVALID, "sdg_sngan_sigmas: null pointer");
  SDG_REQUIRE(c->loaded && c->arch != SDG_ARCH_DCGAN32, SDG_E_STATE, "sdg_sngan_sigmas: no SNGAN weights loaded");
  SDG_CUDA(cudaMemcpyAsync(sigma_out, c->sigma.p, sizeof(float) * c->n_layers, cudaMemcpyDeviceToDevice,
                           (cudaStream_t)stream));
  return 0;
}

static const int kDcganSpec[6][3] = {{3, 16, 2}, {16, 32, 1}, {32, 64, 2}, {64, 128, 1}, {128, 256, 2}, {256, 512, 1}};

extern "C" int sdg_dcgan_load(sdg_ctx* c, const float* const* conv_w, const float* const* bn_gamma,
                              const float* const* bn_beta, const float* const* bn_mean, const float* const* bn_var,
                              const float* fc_w, const float* fc_b, int precision, void* stream) {
  SDG_REQUIRE(c && conv_w && bn_gamma && bn_beta && bn_mean && bn_var && fc_w && fc_b, SDG_E_INVALID,
              "sdg_dcgan_load: null pointer");
  SDG_REQUIRE(precision == SDG_PREC_FP32 || precision == SDG_PREC_BF16 || precision == SDG_PREC_FP16, SDG_E_INVALID,
              "sdg_dcgan_load: precision=%d", precision);
  cudaStream_t s = (cudaStream_t)stream;
  SDG_CUDA(cudaSetDevice(c->device));
  RangeScope range_scope(c->range_flag);
  c->loaded = false;
  c->arch = SDG_ARCH_DCGAN32; c->precision = precision; c->size = 32; c->blocks.clear();
  const bool tc = precision != SDG_PREC_FP32;
  if (tc) { int rc = conv_tc_init(c->device); if (rc) return rc; }
  if (c->convs.size() != 6) {
    for (auto& l : c->convs) l.release_all();
    c->convs.assign(6, ConvLayer());
  }
  { int rc = c->bn_scratch.ensure(sizeof(float) * 512); if (rc) return rc; }
  for (int i = 0; i < 6; ++i) {
    ConvLayer& l = c->convs[i];
    l.cin = kDcganSpec[i][0]; l.cout = kDcganSpec[i][1]; l.stride = kDcganSpec[i][2]; l.ks = 3;
    SDG_REQUIRE(conv_w[i], SDG_E_INVALID, "sdg_dcgan_load: conv %d null", i);
    const float* scale = nullptr;
    l.has_bias = i > 0;
    // tensor-core path: conv 1 stays an fp32 CUDA-core kernel (K = 27); convs 2..6 run on tcgen05 with their channels padded
    // to whole 64-channel TMA chunks (zero weights, zero bias: leaky_relu(0) = 0 keeps the padding channels at zero), the
    // eval-mode BatchNorm folded into weights and bias, and 1/sqrt(2) folded into both so that the FusedLeakyReLU epilogue
    // (leaky_relu(., 0.2) * sqrt(2), positively homogeneous) evaluates the plain LeakyReLU(0.2) of mnist.py:164
    const bool h = tc && i > 0;
    const float mul = h ? 0.70710678118654752f : 1.0f;
    const int cout_p = h ? (l.cout + 63) / 64 * 64 : l.cout, cin_p = h ? (l.cin + 63) / 64 * 64 : l.cin;
    if (i > 0) {
      SDG_REQUIRE(bn_gamma[i - 1] && bn_beta[i - 1] && bn_mean[i - 1] && bn_var[i - 1], SDG_E_INVALID,
                  "sdg_dcgan_load: bn %d null", i);
      { int rc = l.bias.ensure(sizeof(float) * cout_p); if (rc) return rc; }
      if (cout_p != l.cout) SDG_CUDA(cudaMemsetAsync(l.bias.p, 0, sizeof(float) * cout_p, s));
      int rc = bn_fold(bn_gamma[i - 1], bn_beta[i - 1], bn_mean[i - 1], bn_var[i - 1], 1e-5f,
                       c->bn_scratch.as<float>(), l.bias.as<float>(), l.cout, s, mul);
      if (rc) return rc;
      scale = c->bn_scratch.as<float>();
    }
    if (h) {
      l.kpad = l.ktot = 9 * cin_p;
      { int rc = l.w16.ensure(sizeof(h16) * (size_t)l.ktot * cout_p); if (rc) return rc; }
      if (cout_p != l.cout) SDG_CUDA(cudaMemsetAsync(l.w16.p, 0, sizeof(h16) * (size_t)l.ktot * cout_p, s));
      int rc = pack_conv_h16(conv_w[i], nullptr, scale, l.w16.as<h16>(), l.cout, cin_p, l.kpad, 3, precision == SDG_PREC_FP16,
                             l.ktot, 0, s, 1.0f, l.cin);
      if (rc) return rc;
    } else {
      { int rc = l.w32.ensure(sizeof(float) * 9 * l.cin * l.cout); if (rc) return rc; }
      int rc = pack_conv_fp32(conv_w[i], nullptr, scale, l.w32.as<float>(), l.cout, l.cin, 3, s);
      if (rc) return rc;
    }
  }
  c->head_len = 8192;
  { int rc = c->head_w.ensure(sizeof(float) * 8192); if (rc) return rc; }
  { int rc = c->head_b.ensure(sizeof(float)); if (rc) return rc; }
  { int rc = permute_fc(fc_w, c->head_w.as<float>(), 512, 16, s); if (rc) return rc; }
  SDG_CUDA(cudaMemcpyAsync(c->head_b.p, fc_b, sizeof(float), cudaMemcpyDeviceToDevice, s));
  c->loaded = true;
  return 0;
}

// ------------------------------------------------------------------------------------------------
static int forward_sngan_fp32(sdg_ctx* c, const void* x, int layout, int64_t nb, float* logits, cudaStream_t s) {
  const int S = c->size;
  float* X = c->xin.as<float>();
  float* PX = c->xpool.as<float>();
  float* h = c->buf[0].as<float>();
  float* f1 = c->buf[1].as<float>();
  float* f2 = c->buf[2].as<float>();
  int rc;
  if ((rc = prep_input_fp32(x, layout, X, nb, S, S, s))) return rc;
  int hw = S;
  for (size_t bi = 0; bi < c->blocks.size(); ++bi) {
    const BlockSpec& bl = c->blocks[bi];
    const ConvLayer& c1 = c->convs[c->block_first_conv[bi]];
    const ConvLayer& c2 = c->convs[c->block_first_conv[bi] + 1];
    const int ho = bl.down ? hw / 2 : hw;
    if (bl.kind == 0) {
      const ConvLayer& sc = c->convs[c->block_first_conv[bi] + 2];
      if ((rc = conv_fp32(X, c1.w32.as<float>(), c1.bias.as<float>(), f1, nb, hw, hw, c1.cin, c1.cout, 3, 1, ACT_NONE, ACT_RELU, s))) return rc;
      if ((rc = conv_fp32(f1, c2.w32.as<float>(), c2.bias.as<float>(), f2, nb, hw, hw, c2.cin, c2.cout, 3, 1, ACT_NONE, ACT_NONE, s))) return rc;
      if ((rc = combine_fp32(X, 1, nullptr, 0, 0, PX, nb, ho, ho, 3, s))) return rc;                   // avg_pool2d(x)
      if ((rc = conv_fp32(PX, sc.w32.as<float>(), sc.bias.as<float>(), f1, nb, ho, ho, sc.cin, sc.cout, 1, 1, ACT_NONE, ACT_NONE, s))) return rc;
      if ((rc = combine_fp32(f2, 1, f1, 0, 0, h, nb, ho, ho, bl.cout, s))) return rc;
    } else {
      if ((rc = conv_fp32(h, c1.w32.as<float>(), c1.bias.as<float>(), f1, nb, hw, hw, c1.cin, c1.cout, 3, 1, ACT_RELU, ACT_RELU, s))) return rc;
      if ((rc = conv_fp32(f1, c2.w32.as<float>(), c2.bias.as<float>(), f2, nb, hw, hw, c2.cin, c2.cout, 3, 1, ACT_NONE, ACT_NONE, s))) return rc;
      if (c->block_has_sc[bi]) {
        const ConvLayer& sc = c->convs[c->block_first_conv[bi] + 2];
        if ((rc = conv_fp32(h, sc.w32.as<float>(), sc.bias.as<float>(), f1, nb, hw, hw, sc.cin, sc.cout, 1, 1,
                            c->inplace_relu ? ACT_RELU : ACT_NONE, ACT_NONE, s))) return rc;
        if ((rc = combine_fp32(f2, bl.down, f1, bl.down, 0, h, nb, ho, ho, bl.cout, s))) return rc;
      } else {
        if ((rc = combine_fp32(f2, 0, h, 0, c->inplace_relu, f1, nb, ho, ho, bl.cout, s))) return rc;
        float* t = h; h = f1; f1 = t;
      }
    }
    hw = ho;
  }
  return head_sumpool_fp32(h, c->head_w.as<float>(), c->head_b.as<float>(), logits, nb, hw * hw, c->head_len, 1, s);
}

static int forward_sngan_h16(sdg_ctx* c, const void* x, int layout, int64_t nb, float* logits, cudaStream_t s) {
  const int f16 = c->precision == SDG_PREC_FP16;
  const int S = c->size;
  const int nblk = (int)c->blocks.size();
  h16* T = c->buf[0].as<h16>();                                   // relu(c1(.)) of the current block, full resolution
  h16* hR[2] = {c->buf[1].as<h16>(), c->buf[2].as<h16>()};        // relu(h): operand of the next conv
  h16* hW[2] = {c->buf[3].as<h16>(), c->buf[4].as<h16>()};        // raw h, 16-bit: operand of a 1x1 shortcut (textbook mode)
  float* hF[2] = {c->buf[5].as<float>(), c->buf[6].as<float>()};  // raw h, fp32: residual stream / head input
  int rc;
  int hw = S, cur = 0;
  bool fused_head_done = false;
  for (int bi = 0; bi < nblk; ++bi) {
    const BlockSpec& bl = c->blocks[bi];
    const int i1 = c->block_first_conv[bi];
    const ConvLayer& c1 = c->convs[i1];
    const ConvLayer& c2 = c->convs[i1 + 1];
    const bool has_sc = c->block_has_sc[bi] != 0;
    const int ho = bl.down ? hw / 2 : hw;
    const bool last = bi + 1 == nblk;
    const bool next_sc = !last && c->block_has_sc[bi + 1];
    const int o = bi == 0 ? 0 : cur ^ 1;
    TcConv a2;
    a2.n = nb; a2.H = hw; a2.W = hw; a2.Cin = c2.cin; a2.Cout = c2.cout; a2.taps = 9;
    a2.in = T; a2.wb = c2.w16.as<h16>(); a2.bias = c2.bias_sum.as<float>();
    a2.pool = bl.down;
    a2.pool4 = c2.pool4;
    // mimicry's in-place ReLU (the default) makes an identity shortcut add relu(h), which is exactly the 16-bit tensor the
    // next conv reads anyway: blocks with an identity shortcut then take their residual from it and NO fp32 copy of a block
    // output is needed (half the epilogue bytes of those layers).  Only the role-swapped kernel implements it.
    static const int res16_env = getenv("SDG_RES16") ? atoi(getenv("SDG_RES16")) : 1;
    const bool res16 = res16_env && c->inplace_relu && conv_tc_swap_active();
    const bool next_identity_res16 = !last && !next_sc && res16 && c->blocks[bi + 1].cout == 128 &&
                                     ((int64_t)nb * ho * ho) % 32 == 0;
    a2.out_relu = last ? nullptr : hR[o];
    a2.out_raw = (next_sc && !c->inplace_relu) ? hW[o] : nullptr;
    a2.out_f32 = (last || (!next_sc && !next_identity_res16)) ? hF[o] : nullptr;
    // last block of SNGAN-32 (8x8, 128 channels, identity shortcut): ReLU -> sum-pool -> SNLinear fused into the epilogue
    static const int fuse_head_env = getenv("SDG_FUSE_HEAD") ? atoi(getenv("SDG_FUSE_HEAD")) : 1;
    const bool fuse_head = fuse_head_env && last && !bl.down && c2.cout == 128 && (ho * ho == 64 || ho * ho == 32) && bl.kind == 1;
    if (fuse_head) {
      SDG_CUDA(cudaMemsetAsync(logits, 0, sizeof(float) * (size_t)nb, s));
      a2.out_f32 = nullptr;
      a2.head_w = c->head_w.as<float>(); a2.head_b = c->head_b.as<float>(); a2.head_out = logits;
      fused_head_done = true;
    }
    // Block 1 of SNGAN-32 / SNGAN-64 as ONE kernel (conv_b1fused.cu): relu(c1(x)) stays in shared memory.  Needs the byte
    // dataset, the 4x4 stride-2 form of c2 and a consumer that only reads relu(h) (mimicry's in-place ReLU; the next block has
    // a shortcut conv).  (SDG_FUSE_B1: 0 = never, 1 = both, 32 / 64 = that architecture only.)
    static const int fuse_b1_env = getenv("SDG_FUSE_B1") ? atoi(getenv("SDG_FUSE_B1")) : 1;
    const bool fuse_b1_on = fuse_b1_env == 1 || fuse_b1_env == S;
    if (bl.kind == 0 && fuse_b1_on && ((S == 32 && c2.cout == 128) || (S == 64 && c2.cout == 64)) && c1.cout == c2.cout &&
        c2.pool4 && c2.w16f.p && layout == SDG_LAYOUT_U8_NHWC && a2.out_relu && !a2.out_raw && !a2.out_f32 && !a2.head_out &&
        conv_tc_swap_active()) {
      if ((rc = prof_begin(c, s))) return rc;
      if ((rc = b1_fused(x, c1.w16.as<h16>(), c1.bias.as<float>(), c2.w16f.as<h16>(), a2.out_relu, nullptr, nb, c2.cout, f16, s)))
        return rc;
      // useful FLOPs of the launch: c2 in the 4x4 stride-2 form (16 taps per pooled pixel) + c1 (27 MACs per pixel and channel)
      if ((rc = prof_end(c, s, 2.0 * (double)nb * hw * hw * c2.cout * (16.0 * c2.cin * 0.25 + 27.0)))) return rc;
    } else if (bl.kind == 0) {
      // DBlockOptimized: c1 straight from the image bytes; shortcut c_sc(avg_pool2d(x)) as 3 FMAs in c2's epilogue
      if (c1.superpix) {
        if ((rc = first_conv(x, layout, c1.w16s.as<h16>(), c1.bias.as<float>(), T, nb, S, c1.cout, f16, s, 1))) return rc;
      } else {
        if ((rc = first_conv(x, layout, c1.w16.as<h16>(), c1.bias.as<float>(), T, nb, S, c1.cout, f16, s))) return rc;
      }
      a2.img = x; a2.img_layout = layout; a2.sc_w3 = c->convs[i1 + 2].w3.as<float>();
      // executed FLOPs of this launch: the 4x4 stride-2 form does 16 taps per pooled pixel instead of 9 per input pixel
      double macs_per_out = c2.pool4 ? 16.0 * c2.cin * 0.25 : 9.0 * c2.cin;
      if (c2.superpix && conv_tc_swap_active()) {
        // two pooled pixels per GEMM pixel: [n, ho, ho/2] grid, 4 x 6 taps at x-stride 4, 128 "channels"
        a2.general = 1; a2.pool = 0; a2.pool4 = 0; a2.img_up = 1; a2.superpix = 1;
        a2.H = ho; a2.W = ho / 2; a2.grid_w = ho / 2; a2.in_H = hw; a2.in_W = hw;
        a2.taps_y = 4; a2.taps_x = 6; a2.sy = 2; a2.sx = 4; a2.offy = -1; a2.offx = -1;
        a2.Cout = 128; a2.wb = c2.w16s.as<h16>(); a2.bias = c2.bias2.as<float>(); a2.sc_w3 = c->convs[i1 + 2].w3s.as<float>();
        macs_per_out = 24.0 * c2.cin * 0.25;            // per original-resolution pixel per output channel, zeros included
      }
      if ((rc = prof_begin(c, s))) return rc;
      if ((rc = conv_tc(a2, f16, s))) return rc;
      if ((rc = prof_end(c, s, 2.0 * (double)nb * hw * hw * c2.cout * macs_per_out))) return rc;
    } else {
      TcConv a1;
      a1.n = nb; a1.H = hw; a1.W = hw; a1.Cin = c1.cin; a1.Cout = c1.cout; a1.taps = 9;
      a1.in = hR[cur]; a1.wb = c1.w16.as<h16>(); a1.bias = c1.bias.as<float>(); a1.out_relu = T;
      if (c1.superpix && conv_tc_swap_active()) {
        // two adjacent output pixels per GEMM pixel: [n, hw, hw/2] grid, 3 x 4 taps at x-stride 2, 128 "channels"
        a1.general = 1; a1.H = hw; a1.W = hw / 2; a1.grid_w = hw / 2; a1.in_H = hw; a1.in_W = hw;
        a1.taps_y = 3; a1.taps_x = 4; a1.sy = 1; a1.sx = 2; a1.offy = -1; a1.offx = -1;
        a1.Cout = 128; a1.wb = c1.w16s.as<h16>(); a1.bias = c1.bias2.as<float>();
      }
      if ((rc = conv_tc(a1, f16, s))) return rc;
      if (has_sc) {                      // 1x1 shortcut conv folded into c2's K loop (input: relu(h) or h)
        a2.sc_in = c->inplace_relu ? hR[cur] : hW[cur];
        a2.sc_C = c->convs[i1 + 2].kpad;
      } else if (res16 && c2.cout == 128 && ((int64_t)nb * hw * hw) % 32 == 0) {
        a2.res_h16 = hR[cur];            // identity shortcut = relu(h), already stored as the 16-bit conv operand
        a2.res_relu = 0;
      } else {                           // identity shortcut from the fp32 residual stream
        a2.res_f32 = hF[cur];
        a2.res_relu = c->inplace_relu;
      }
      if ((rc = conv_tc(a2, f16, s))) return rc;
    }
    cur = o;
    hw = ho;
  }
  if (fused_head_done) return 0;
  return head_sumpool_fp32(hF[cur], c->head_w.as<float>(), c->head_b.as<float>(), logits, nb, hw * hw, c->head_len, 1, s);
}

static int forward_dcgan_fp32(sdg_ctx* c, const void* x, int layout, int64_t nb, float* logits, cudaStream_t s) {
  float* X = c->xin.as<float>();
  float* a = c->buf[0].as<float>();
  float* b = c->buf[1].as<float>();
  int rc;
  if ((rc = prep_input_fp32(x, layout, X, nb, 32, 32, s))) return rc;
  const float* in = X;
  int hw = 32;
  for (int i = 0; i < 6; ++i) {
    const ConvLayer& l = c->convs[i];
    if ((rc = conv_fp32(in, l.w32.as<float>(), l.has_bias ? l.bias.as<float>() : nullptr, a, nb, hw, hw, l.cin, l.cout,
                        3, l.stride, ACT_NONE, ACT_LRELU, s))) return rc;
    hw = (hw + 2 - 3) / l.stride + 1;
    in = a;
    float* t = a; a = b; b = t;
  }
  return head_dot_fp32(in, c->head_w.as<float>(), c->head_b.as<float>(), logits, nb, c->head_len, s);
}

// DCGAN discriminator on the tensor cores (mnist.py:161-192, eval mode): conv 1 from the image bytes on CUDA cores, convs 2..6
// as tcgen05 implicit GEMMs (stride-2 layers through the TMA traversal stride, BatchNorm + LeakyReLU in the epilogue), the
// 8192 -> 1 Linear head as an fp32 dot product over the fp32 output of conv 6.
static int forward_dcgan_h16(sdg_ctx* c, const void* x, int layout, int64_t nb, float* logits, cudaStream_t s) {
  const int f16 = c->precision == SDG_PREC_FP16;
  h16* a = c->buf[0].as<h16>();
  h16* b = c->buf[1].as<h16>();
  float* f = c->buf[2].as<float>();
  int rc;
  if ((rc = dcgan_first_conv_h16(x, layout, c->convs[0].w32.as<float>(), a, nb, 32, f16, s))) return rc;
  int hw = 16;
  for (int i = 1; i < 6; ++i) {
    const ConvLayer& l = c->convs[i];
    const int cin_p = (l.cin + 63) / 64 * 64, cout_p = (l.cout + 63) / 64 * 64;
    const int ho = l.stride == 2 ? hw / 2 : hw;
    TcConv t;
    t.n = nb; t.H = ho; t.W = ho; t.Cin = cin_p; t.Cout = cout_p; t.taps = 9;
    t.in = a; t.wb = l.w16.as<h16>(); t.bias = l.bias.as<float>(); t.act = 1;
    if (l.stride == 2) { t.stride = 2; t.in_H = hw; t.in_W = hw; }
    if (i == 5) t.out_f32 = f; else t.out_raw = b;
    if (i == 5 && (rc = prof_begin(c, s))) return rc;
    if ((rc = conv_tc(t, f16, s))) return rc;
    if (i == 5 && (rc = prof_end(c, s, 2.0 * (double)nb * ho * ho * l.cout * 9.0 * l.cin))) return rc;
    hw = ho;
    h16* tmp = a; a = b; b = tmp;
  }
  return head_dot_fp32(f, c->head_w.as<float>(), c->head_b.as<float>(), logits, nb, c->head_len, s);
}

extern "C" int sdg_d_forward(sdg_ctx* c, const void* x, int layout, int64_t n, float* logits_out, void* stream) {
  SDG_REQUIRE(c, SDG_E_INVALID, "sdg_d_forward: null context");
  SDG_REQUIRE(c->loaded, SDG_E_STATE, "sdg_d_forward: no discriminator weights loaded");
  SDG_REQUIRE(layout == SDG_LAYOUT_U8_NHWC || layout == SDG_LAYOUT_F32_NCHW, SDG_E_INVALID, "sdg_d_forward: layout=%d", layout);
  SDG_REQUIRE(n >= 0, SDG_E_INVALID, "sdg_d_forward: n=%lld", (long long)n);
  if (n == 0) return 0;          // an empty shard (world > N, StyleGAN2 drop_last tail): torch hands out null pointers for it
  SDG_REQUIRE(x && logits_out, SDG_E_INVALID, "sdg_d_forward: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  SDG_CUDA(cudaSetDevice(c->device));
  RangeScope range_scope(c->range_flag);
  const int S = c->size;
  const bool bf = c->precision != SDG_PREC_FP32;
  if (c->arch == SDG_ARCH_STYLEGAN2) {
    const int B = c->sg2_batch;
    SDG_REQUIRE(n % B == 0, SDG_E_INVALID, "sdg_d_forward: StyleGAN2 needs whole reference batches (minibatch-stddev): n=%lld "
                "is not a multiple of batch %d; drop the ragged tail like the reference's drop_last=True", (long long)n, B);
    const bool tc = c->precision != SDG_PREC_FP32;
    const int64_t el = sg2_buf_elems(c);
    int64_t chunk = c->chunk;
    if (chunk <= 0) chunk = tc ? (8LL << 30) / sg2_h16_bytes_per_sample(c) : (3LL << 30) / (4 * el * 4 + (int64_t)S * S * 12);
    chunk = std::max<int64_t>(B, chunk / B * B);
    if (chunk > n) chunk = n;
    if (tc) {
      int rc = sg2_ensure_h16(c, chunk);
      if (rc) return rc;
    } else {
      for (int i = 0; i < 4; ++i) { int rc = c->buf[i].ensure((size_t)chunk * el * 4); if (rc) return rc; }
      { int rc = c->xin.ensure((size_t)chunk * S * S * 3 * 4); if (rc) return rc; }
    }
    { int rc = c->sg2_sd.ensure(sizeof(float) * (size_t)chunk); if (rc) return rc; }
    const size_t in_stride = layout == SDG_LAYOUT_U8_NHWC ? (size_t)S * S * 3 : (size_t)S * S * 3 * 4;
    for (int64_t s0 = 0; s0 < n; s0 += chunk) {
      const int64_t nb = (n - s0) < chunk ? (n - s0) : chunk;
      const void* xs = (const char*)x + (size_t)s0 * in_stride;
      int rc = tc ? forward_stylegan2_h16(c, xs, layout, nb, logits_out + s0, s)
                  : forward_stylegan2_fp32(c, xs, layout, nb, logits_out + s0, s);
      if (rc) return rc;
    }
    return 0;
  }
  const int64_t act = max_act_elems(c);
  int64_t chunk = c->chunk;
  if (chunk <= 0) {
    // default: about 4 GiB for the largest activation buffer (~10 GB of scratch in all; measured on the SNGAN-32 pass:
    // 2048 samples per sweep 19.95 ms, 4096 19.45, 16384 18.53), sweeps balanced so that no small remainder is left
    const int64_t bytes_per = act * (bf ? 2 : 4);
    chunk = (4LL << 30) / bytes_per;
    if (chunk < 1) chunk = 1;
    if (chunk < n) {
      const int64_t sweeps = cdiv(n, chunk);
      chunk = (cdiv(n, sweeps) + 7) / 8 * 8;
    }
  }
  if (chunk > n) chunk = n;
  if (bf && c->arch == SDG_ARCH_DCGAN32) {
    // two 16-bit ping-pong buffers of [16,16,64] per sample, one fp32 [4,4,512]
    for (int i = 0; i < 2; ++i) { int rc = c->buf[i].ensure((size_t)chunk * 16 * 16 * 64 * 2); if (rc) return rc; }
    { int rc = c->buf[2].ensure((size_t)chunk * 16 * 512 * 4); if (rc) return rc; }
  } else if (bf) {
    // T: full-resolution relu(c1) of any block; hR/hW: block outputs 16-bit; hF: block outputs fp32
    int64_t t_el = 0, h_el = 0;
    int hw = S;
    for (auto& b : c->blocks) {
      const int hidden = b.kind == 0 ? b.cout : b.cin;
      const int ho = b.down ? hw / 2 : hw;
      t_el = std::max<int64_t>(t_el, (int64_t)hw * hw * hidden);
      h_el = std::max<int64_t>(h_el, (int64_t)ho * ho * b.cout);
      hw = ho;
    }
    { int rc = c->buf[0].ensure((size_t)chunk * t_el * 2); if (rc) return rc; }
    for (int i = 1; i <= 2; ++i) { int rc = c->buf[i].ensure((size_t)chunk * h_el * 2); if (rc) return rc; }
    if (!c->inplace_relu) for (int i = 3; i <= 4; ++i) { int rc = c->buf[i].ensure((size_t)chunk * h_el * 2); if (rc) return rc; }
    for (int i = 5; i <= 6; ++i) { int rc = c->buf[i].ensure((size_t)chunk * h_el * 4); if (rc) return rc; }
  } else {
    const int nbuf = c->arch == SDG_ARCH_DCGAN32 ? 2 : 3;
    for (int i = 0; i < nbuf; ++i) { int rc = c->buf[i].ensure((size_t)chunk * act * 4); if (rc) return rc; }
    { int rc = c->xin.ensure((size_t)chunk * S * S * 3 * 4); if (rc) return rc; }
    { int rc = c->xpool.ensure((size_t)chunk * (S / 2) * (S / 2) * 3 * 4); if (rc) return rc; }
  }
  const size_t in_stride = layout == SDG_LAYOUT_U8_NHWC ? (size_t)S * S * 3 : (size_t)S * S * 3 * 4;
  for (int64_t s0 = 0; s0 < n; s0 += chunk) {
    const int64_t nb = (n - s0) < chunk ? (n - s0) : chunk;
    const void* xs = (const char*)x + (size_t)s0 * in_stride;
    int rc;
    if (c->arch == SDG_ARCH_DCGAN32) rc = bf ? forward_dcgan_h16(c, xs, layout, nb, logits_out + s0, s)
                                             : forward_dcgan_fp32(c, xs, layout, nb, logits_out + s0, s);
    else if (bf) rc = forward_sngan_h16(c, xs, layout, nb, logits_out + s0, s);
    else rc = forward_sngan_fp32(c, xs, layout, nb, logits_out + s0, s);
    if (rc) return rc;
  }
  return 0;
}

extern "C" int sdg_conv2d_h16(const void* in, const void* wb, const float* bias, int64_t n, int H, int W, int Cin,
                              int Cout, int ks, const void* sc_in, int sc_C, int pool, const float* res_f32, int res_relu,
                              const void* img, int img_layout, const float* sc_w3, void* out_relu, void* out_raw,
                              float* out_f32, int precision, void* stream) {
  SDG_REQUIRE(in && wb, SDG_E_INVALID, "sdg_conv2d_h16: null pointer");
  SDG_REQUIRE(ks == 1 || ks == 3, SDG_E_UNSUPPORTED, "sdg_conv2d_h16: ks=%d", ks);
  SDG_REQUIRE(precision == SDG_PREC_BF16 || precision == SDG_PREC_FP16, SDG_E_INVALID, "sdg_conv2d_h16: precision=%d",
              precision);
  int dev = 0;
  SDG_CUDA(cudaGetDevice(&dev));
  { int rc = conv_tc_init(dev); if (rc) return rc; }
  TcConv a;
  a.in = (const h16*)in; a.wb = (const h16*)wb; a.bias = bias; a.n = n; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout;
  a.taps = ks * ks; a.sc_in = (const h16*)sc_in; a.sc_C = sc_C; a.pool = pool != 0; a.pool4 = pool == 2;
  a.res_f32 = res_f32; a.res_relu = res_relu;
  a.img = img; a.img_layout = img_layout; a.sc_w3 = sc_w3;
  a.out_relu = (h16*)out_relu; a.out_raw = (h16*)out_raw; a.out_f32 = out_f32;
  return conv_tc(a, precision == SDG_PREC_FP16, (cudaStream_t)stream);
}

extern "C" int sdg_conv2d_sg2_h16(const void* in, const void* wb, const float* bias, int64_t n, int Hout, int Wout, int in_H,
                                  int in_W, int Cin, int Cout, int ks, int stride, int pad, int act, const void* skip_in,
                                  int skip_C, const float* res_f32, float out_scale, void* out_raw, float* out_f32,
                                  int precision, void* stream) {
  SDG_REQUIRE(in && wb, SDG_E_INVALID, "sdg_conv2d_sg2_h16: null pointer");
  SDG_REQUIRE(ks == 1 || ks == 3, SDG_E_UNSUPPORTED, "sdg_conv2d_sg2_h16: ks=%d", ks);
  SDG_REQUIRE(precision == SDG_PREC_BF16 || precision == SDG_PREC_FP16, SDG_E_INVALID, "sdg_conv2d_sg2_h16: precision=%d",
              precision);
  int dev = 0;
  SDG_CUDA(cudaGetDevice(&dev));
  { int rc = conv_tc_init(dev); if (rc) return rc; }
  TcConv a;
  a.in = (const h16*)in; a.wb = (const h16*)wb; a.bias = bias; a.n = n; a.H = Hout; a.W = Wout; a.in_H = in_H; a.in_W = in_W;
  a.Cin = Cin; a.Cout = Cout; a.taps = ks * ks; a.stride = stride; a.no_pad = pad ? 0 : 1; a.act = act;
  a.sc_in = (const h16*)skip_in; a.sc_C = skip_C; a.sc_sep = skip_in ? 1 : 0;
  a.res_f32 = res_f32; a.out_scale = out_scale; a.out_raw = (h16*)out_raw; a.out_f32 = out_f32;
  return conv_tc(a, precision == SDG_PREC_FP16, (cudaStream_t)stream);
}

static int block1_fused_entry(const char* who, int ch, const void* x, const void* w1, const float* b1, const void* w2,
                              const float* bias2, const float* sc_w3, void* out_relu, void* dbg_t, int64_t n, int precision,
                              void* stream) {
  SDG_REQUIRE(precision == SDG_PREC_BF16 || precision == SDG_PREC_FP16, SDG_E_INVALID, "%s: precision=%d", who, precision);
  SDG_REQUIRE(n >= 0, SDG_E_INVALID, "%s: n=%lld", who, (long long)n);
  int dev = 0;
  SDG_CUDA(cudaGetDevice(&dev));
  { int rc = conv_tc_init(dev); if (rc) return rc; }
  SDG_REQUIRE(w2 && bias2 && sc_w3, SDG_E_INVALID, "%s: null pointer", who);
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int f16 = precision == SDG_PREC_FP16;
  h16* w2f = nullptr;
  SDG_CUDA(cudaMallocAsync((void**)&w2f, sizeof(h16) * (size_t)b1_fused_w2_elems(ch), st));
  int rc = b1_fused_pack((const h16*)w2, sc_w3, bias2, w2f, ch, f16, st);
  if (!rc) rc = b1_fused(x, (const h16*)w1, b1, w2f, (h16*)out_relu, (h16*)dbg_t, n, ch, f16, st);
  cudaFreeAsync(w2f, st);
  return rc;
}

extern "C" int sdg_sngan32_block1_fused_h16(const void* x, const void* w1, const float* b1, const void* w2, const float* bias2,
                                            const float* sc_w3, void* out_relu, void* dbg_t, int64_t n, int precision,
                                            void* stream) {
  return block1_fused_entry("sdg_sngan32_block1_fused_h16", 128, x, w1, b1, w2, bias2, sc_w3, out_relu, dbg_t, n, precision, stream);
}

extern "C" int sdg_sngan64_block1_fused_h16(const void* x, const void* w1, const float* b1, const void* w2, const float* bias2,
                                            const float* sc_w3, void* out_relu, void* dbg_t, int64_t n, int precision,
                                            void* stream) {
  return block1_fused_entry("sdg_sngan64_block1_fused_h16", 64, x, w1, b1, w2, bias2, sc_w3, out_relu, dbg_t, n, precision, stream);
}

extern "C" int sdg_blur_h16(const void* in, void* out, int64_t n, int H, int W, int C, int pad, int stride, int precision,
                            void* stream) {
  SDG_REQUIRE(in && out, SDG_E_INVALID, "sdg_blur_h16: null pointer");
  SDG_REQUIRE(precision == SDG_PREC_BF16 || precision == SDG_PREC_FP16, SDG_E_INVALID, "sdg_blur_h16: precision=%d", precision);
  SDG_REQUIRE(C % 8 == 0 && (stride == 1 || stride == 2) && pad >= 0 && pad <= 3 && H + 2 * pad >= 4 && W + 2 * pad >= 4,
              SDG_E_INVALID, "sdg_blur_h16: C=%d pad=%d stride=%d H=%d W=%d", C, pad, stride, H, W);
  int dev = 0;
  SDG_CUDA(cudaGetDevice(&dev));
  { int rc = conv_tc_init(dev); if (rc) return rc; }      // the TMA-fed blur needs the tensor-map encoder
  return blur_h16((const h16*)in, (h16*)out, n, H, W, C, pad, stride, precision == SDG_PREC_FP16, (cudaStream_t)stream);
}

extern "C" int sdg_first_conv_h16(const void* x, int layout, const void* wb, const float* bias, void* out, int64_t n,
                                  int S, int Cout, int precision, void* stream) {
  SDG_REQUIRE(x && wb && bias && out, SDG_E_INVALID, "sdg_first_conv_h16: null pointer");
  SDG_REQUIRE(precision == SDG_PREC_BF16 || precision == SDG_PREC_FP16, SDG_E_INVALID, "sdg_first_conv_h16: precision=%d",
              precision);
  SDG_REQUIRE(layout == SDG_LAYOUT_U8_NHWC || layout == SDG_LAYOUT_F32_NCHW, SDG_E_INVALID, "sdg_first_conv_h16: layout=%d",
              layout);
  int dev = 0;
  SDG_CUDA(cudaGetDevice(&dev));
  { int rc = conv_tc_init(dev); if (rc) return rc; }
  return first_conv(x, layout, (const h16*)wb, bias, (h16*)out, n, S, Cout, precision == SDG_PREC_FP16, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// StyleGAN2 discriminator (diagan-pkg/diagan/models/stylegan2.py:619-677), fp32 CUDA-core path
// ------------------------------------------------------------------------------------------------
static int sg2_channels(int res) {
  switch (res) {
    case 4: case 8: case 16: case 32: case 64: return 512;
    case 128: return 256;
    case 256: return 128;
    case 512: return 64;
    case 1024: return 32;
  }
  return 0;
}

extern "C" int sdg_ctx_set_batch(sdg_ctx* c, int batch) {
  SDG_REQUIRE(c && batch >= 1, SDG_E_INVALID, "sdg_ctx_set_batch: bad argument");
  const int group = batch < 4 ? batch : 4;
  SDG_REQUIRE(batch % group == 0, SDG_E_INVALID, "sdg_ctx_set_batch: batch %d is not divisible by the stddev group %d (the "
              "reference's view() fails for it too)", batch, group);
  c->sg2_batch = batch;
  return 0;
}

extern "C" int sdg_stylegan2_load(sdg_ctx* c, int size, int n_tensors, const float* const* t, int precision, void* stream) {
  SDG_REQUIRE(c && t, SDG_E_INVALID, "sdg_stylegan2_load: null pointer");
  SDG_REQUIRE(precision == SDG_PREC_FP32 || precision == SDG_PREC_BF16 || precision == SDG_PREC_FP16, SDG_E_INVALID,
              "sdg_stylegan2_load: precision=%d", precision);
  SDG_REQUIRE(size >= 8 && size <= 1024 && (size & (size - 1)) == 0, SDG_E_INVALID, "sdg_stylegan2_load: size=%d", size);
  cudaStream_t s = (cudaStream_t)stream;
  SDG_CUDA(cudaSetDevice(c->device));
  RangeScope range_scope(c->range_flag);
  c->loaded = false;
  c->sg2_blocks.clear();
  int cin = sg2_channels(size);
  for (int res = size; res > 4; res >>= 1) {
    const int cout = sg2_channels(res >> 1);
    c->sg2_blocks.push_back({cin, cout});
    cin = cout;
  }
  const int nblk = (int)c->sg2_blocks.size();
  const int want = 2 + 5 * nblk + 6;
  SDG_REQUIRE(n_tensors == want, SDG_E_INVALID, "sdg_stylegan2_load: size %d needs %d tensors, got %d", size, want, n_tensors);
  for (int i = 0; i < n_tensors; ++i) SDG_REQUIRE(t[i], SDG_E_INVALID, "sdg_stylegan2_load: tensor %d is null", i);
  c->arch = SDG_ARCH_STYLEGAN2; c->precision = precision; c->size = size; c->blocks.clear();
  const int n_convs = 1 + 3 * nblk + 2;                  // first, (conv1, conv2, skip) per block, final conv, linear 0
  if ((int)c->convs.size() != n_convs) {
    for (auto& l : c->convs) l.release_all();
    c->convs.assign(n_convs, ConvLayer());
  }
  const bool tc = precision != SDG_PREC_FP32;
  if (tc) { int rc = conv_tc_init(c->device); if (rc) return rc; }
  // equalised learning rate: the run-time scale 1/sqrt(Cin*k*k) (stylegan2.py:103,117) is folded into the packed weights.
  // Tensor-core mode packs the ResBlock convs as 16-bit [Cout][tap*Cin + c]; the 3-channel first conv and the tail
  // (513-channel conv on 4x4, the two EqualLinear: 0.1 % of the FLOPs) stay fp32.
  auto pack = [&](ConvLayer& l, const float* W, const float* b, int cout, int cin_, int ks, bool h) -> int {
    l.cout = cout; l.cin = cin_; l.ks = ks; l.has_bias = b != nullptr;
    const float mul = 1.0f / sqrtf((float)cin_ * ks * ks);
    if (h) {
      l.kpad = l.ktot = ks * ks * cin_;
      { int rc = l.w16.ensure(sizeof(h16) * (size_t)l.ktot * cout); if (rc) return rc; }
      { int rc = pack_conv_h16(W, nullptr, nullptr, l.w16.as<h16>(), cout, cin_, l.kpad, ks, precision == SDG_PREC_FP16, l.ktot, 0,
                               s, mul);
        if (rc) return rc; }
    } else {
      { int rc = l.w32.ensure(sizeof(float) * (size_t)ks * ks * cin_ * cout); if (rc) return rc; }
      { int rc = pack_conv_fp32(W, nullptr, nullptr, l.w32.as<float>(), cout, cin_, ks, s, mul); if (rc) return rc; }
    }
    if (b) {
      { int rc = l.bias.ensure(sizeof(float) * cout); if (rc) return rc; }
      SDG_CUDA(cudaMemcpyAsync(l.bias.p, b, sizeof(float) * cout, cudaMemcpyDeviceToDevice, s));
    }
    return 0;
  };
  int ti = 0, li = 0;
  const int c0 = sg2_channels(size);
  { int rc = pack(c->convs[li++], t[ti], t[ti + 1], c0, 3, 1, false); if (rc) return rc; ti += 2; }
  for (auto& b : c->sg2_blocks) {
    { int rc = pack(c->convs[li++], t[ti], t[ti + 1], b.first, b.first, 3, tc); if (rc) return rc; }
    if (tc) {
      // conv2 and the skip conv share one launch: wb = [Cout][9*Cin (conv2 taps) | Cin (skip 1x1)], two accumulators
      ConvLayer& l2 = c->convs[li++];
      l2.cout = b.second; l2.cin = b.first; l2.ks = 3; l2.has_bias = true;
      l2.kpad = 9 * b.first; l2.ktot = 10 * b.first;
      const int f16 = precision == SDG_PREC_FP16;
      { int rc = l2.w16.ensure(sizeof(h16) * (size_t)l2.ktot * l2.cout); if (rc) return rc; }
      { int rc = pack_conv_h16(t[ti + 2], nullptr, nullptr, l2.w16.as<h16>(), l2.cout, l2.cin, l2.kpad, 3, f16, l2.ktot, 0, s,
                               1.0f / sqrtf(9.f * b.first)); if (rc) return rc; }
      { int rc = pack_conv_h16(t[ti + 4], nullptr, nullptr, l2.w16.as<h16>(), l2.cout, l2.cin, l2.cin, 1, f16, l2.ktot, l2.kpad, s,
                               1.0f / sqrtf((float)b.first)); if (rc) return rc; }
      { int rc = l2.bias.ensure(sizeof(float) * l2.cout); if (rc) return rc; }
      SDG_CUDA(cudaMemcpyAsync(l2.bias.p, t[ti + 3], sizeof(float) * l2.cout, cudaMemcpyDeviceToDevice, s));
      ConvLayer& lk = c->convs[li++];
      lk.cout = b.second; lk.cin = b.first; lk.ks = 1; lk.has_bias = false;
    } else {
      { int rc = pack(c->convs[li++], t[ti + 2], t[ti + 3], b.second, b.first, 3, tc); if (rc) return rc; }
      { int rc = pack(c->convs[li++], t[ti + 4], nullptr, b.second, b.first, 1, tc); if (rc) return rc; }
    }
    ti += 5;
  }
  if (tc) {
    // final_conv: the 512 feature channels on the tensor cores; the minibatch-stddev channel (input channel 512) is constant
    // over the 4x4 map, so its contribution is sd[n] * (sum of its in-bounds tap weights): w3 holds that [16][512] table
    ConvLayer& l = c->convs[li++];
    const float mul = 1.0f / sqrtf(513.f * 9.f);
    // split-precision operands (sg2_fp32.cu): K = 9 taps x [hi | lo | hi] of 512 channels
    l.cout = 512; l.cin = 3 * 512; l.ks = 3; l.has_bias = true; l.kpad = l.ktot = 9 * 3 * 512;
    { int rc = l.w16.ensure(sizeof(h16) * (size_t)l.ktot * 512); if (rc) return rc; }
    { int rc = pack_split3_h16(t[ti], mul, l.w16.as<h16>(), 512, 512, 9, 513, 0, precision == SDG_PREC_FP16, s); if (rc) return rc; }
    { int rc = l.w3.ensure(sizeof(float) * 16 * 512); if (rc) return rc; }
    { int rc = pack_const_channel_fp32(t[ti], mul, l.w3.as<float>(), 512, 513, 512, 4, s); if (rc) return rc; }
    { int rc = l.bias.ensure(sizeof(float) * 512); if (rc) return rc; }
    SDG_CUDA(cudaMemcpyAsync(l.bias.p, t[ti + 1], sizeof(float) * 512, cudaMemcpyDeviceToDevice, s));
    ti += 2;
    ConvLayer& l0 = c->convs[li++];          // EqualLinear(8192, 512, fused_lrelu): a [n][8192] x [512][8192]^T GEMM
    l0.cout = 512; l0.cin = 3 * 8192; l0.ks = 1; l0.has_bias = true; l0.kpad = l0.ktot = 3 * 8192;
    { int rc = l0.w16.ensure(sizeof(h16) * (size_t)l0.ktot * 512); if (rc) return rc; }
    { int rc = pack_split3_h16(t[ti], 1.0f / sqrtf(8192.f), l0.w16.as<h16>(), 512, 512, 16, 0, 1, precision == SDG_PREC_FP16, s);
      if (rc) return rc; }
    { int rc = l0.bias.ensure(sizeof(float) * 512); if (rc) return rc; }
    SDG_CUDA(cudaMemcpyAsync(l0.bias.p, t[ti + 1], sizeof(float) * 512, cudaMemcpyDeviceToDevice, s));
    ti += 2;
  } else {
    { int rc = pack(c->convs[li++], t[ti], t[ti + 1], 512, 513, 3, false); if (rc) return rc; ti += 2; }
    ConvLayer& l = c->convs[li++];           // EqualLinear(8192, 512, fused_lrelu) as a 1x1 conv over the NHWC-flattened map
    l.cout = 512; l.cin = 8192; l.ks = 1; l.has_bias = true;
    { int rc = l.w32.ensure(sizeof(float) * 8192 * 512); if (rc) return rc; }
    { int rc = pack_linear_nchw_fp32(t[ti], 1.0f / sqrtf(8192.f), l.w32.as<float>(), 512, 512, 16, s); if (rc) return rc; }
    { int rc = l.bias.ensure(sizeof(float) * 512); if (rc) return rc; }
    SDG_CUDA(cudaMemcpyAsync(l.bias.p, t[ti + 1], sizeof(float) * 512, cudaMemcpyDeviceToDevice, s));
    ti += 2;
  }
  c->head_len = 512;
  { int rc = c->head_w.ensure(sizeof(float) * 512); if (rc) return rc; }
  { int rc = c->head_b.ensure(sizeof(float)); if (rc) return rc; }
  { int rc = pack_conv_fp32(t[ti], nullptr, nullptr, c->head_w.as<float>(), 1, 512, 1, s, 1.0f / sqrtf(512.f)); if (rc) return rc; }
  SDG_CUDA(cudaMemcpyAsync(c->head_b.p, t[ti + 1], sizeof(float), cudaMemcpyDeviceToDevice, s));
  c->loaded = true;
  return 0;
}

static int64_t sg2_buf_elems(const sdg_ctx* c) {
  int64_t m = (int64_t)c->size * c->size * sg2_channels(c->size);
  int hw = c->size;
  for (auto& b : c->sg2_blocks) {
    const int cm = b.first > b.second ? b.first : b.second;
    m = std::max<int64_t>(m, (int64_t)(hw + 1) * (hw + 1) * cm);
    hw >>= 1;
  }
  return std::max<int64_t>(m, 16 * 513);
}

static int forward_stylegan2_fp32(sdg_ctx* c, const void* x, int layout, int64_t nb, float* logits, cudaStream_t s) {
  const int S = c->size;
  float* X = c->xin.as<float>();
  float* A = c->buf[0].as<float>();
  float* B = c->buf[1].as<float>();
  float* C = c->buf[2].as<float>();
  float* D = c->buf[3].as<float>();
  int rc, li = 0;
  if ((rc = prep_input_fp32(x, layout, X, nb, S, S, s))) return rc;
  {
    const ConvLayer& l = c->convs[li++];
    if ((rc = conv_fp32(X, l.w32.as<float>(), l.bias.as<float>(), A, nb, S, S, 3, l.cout, 1, 1, ACT_NONE, ACT_LRELU_SQRT2, s, 0))) return rc;
  }
  int hw = S;
  for (auto& b : c->sg2_blocks) {
    const ConvLayer& c1 = c->convs[li++];
    const ConvLayer& c2 = c->convs[li++];
    const ConvLayer& sk = c->convs[li++];
    const int ho = hw / 2;
    // conv1: 3x3 + FusedLeakyReLU;  conv2: Blur(pad 2) + 3x3 stride 2 + FusedLeakyReLU;  skip: Blur(pad 1) + 1x1 stride 2
    if ((rc = conv_fp32(A, c1.w32.as<float>(), c1.bias.as<float>(), B, nb, hw, hw, b.first, b.first, 3, 1, ACT_NONE, ACT_LRELU_SQRT2, s, 1))) return rc;
    if ((rc = blur_fp32(B, C, nb, hw, hw, b.first, 2, s))) return rc;
    if ((rc = conv_fp32(C, c2.w32.as<float>(), c2.bias.as<float>(), D, nb, hw + 1, hw + 1, b.first, b.second, 3, 2, ACT_NONE, ACT_LRELU_SQRT2, s, 0))) return rc;
    if ((rc = blur_fp32(A, C, nb, hw, hw, b.first, 1, s))) return rc;
    if ((rc = conv_fp32(C, sk.w32.as<float>(), nullptr, B, nb, hw - 1, hw - 1, b.first, b.second, 1, 2, ACT_NONE, ACT_NONE, s, 0))) return rc;
    if ((rc = add_div_sqrt2_fp32(D, B, A, nb * ho * ho * (int64_t)b.second, s))) return rc;
    hw = ho;
  }
  // minibatch-stddev channel, final conv, the two EqualLinear layers
  if ((rc = minibatch_stddev_cat_fp32(A, B, c->sg2_sd.as<float>(), nb, c->sg2_batch, 16, 512, s))) return rc;
  const ConvLayer& fc = c->convs[li++];
  if ((rc = conv_fp32(B, fc.w32.as<float>(), fc.bias.as<float>(), D, nb, 4, 4, 513, 512, 3, 1, ACT_NONE, ACT_LRELU_SQRT2, s, 1))) return rc;
  const ConvLayer& l0 = c->convs[li++];
  if ((rc = conv_fp32(D, l0.w32.as<float>(), l0.bias.as<float>(), C, nb, 1, 1, 8192, 512, 1, 1, ACT_NONE, ACT_LRELU_SQRT2, s, 0))) return rc;
  return head_dot_fp32(C, c->head_w.as<float>(), c->head_b.as<float>(), logits, nb, 512, s);
}

// ------------------------------------------------------------------------------------------------
// StyleGAN2 discriminator, tensor-core path: ResBlock convs on tcgen05 (conv_tc.cu), activations NHWC 16-bit.
//   conv1  = conv3x3 + FusedLeakyReLU                       -> one conv_tc launch (act in the epilogue)
//   skip   = Blur(pad 1) + conv1x1 stride 2                 -> blur evaluated only at the even outputs (16-bit, output resolution)
//   conv2  = Blur(pad 2) + conv3x3 stride 2 + FusedLeakyReLU -> blur, then ONE strided conv_tc launch that also runs the skip's 1x1
//            conv as extra K iterations into a second TMEM accumulator; epilogue: (flrelu(acc0 + b) + acc1) / sqrt(2)
//            (stylegan2.py:611-614)
// Per-sample scratch (elements): A, B = max hw^2*Cin; C = max (hw+1)^2*Cin; S = max (hw/2)^2*Cin.
struct Sg2Sizes { int64_t a, c, sd, f; };

static Sg2Sizes sg2_h16_sizes(const sdg_ctx* c) {
  Sg2Sizes z = {0, 0, 0, 0};
  int hw = c->size;
  for (auto& b : c->sg2_blocks) {
    const int ho = hw / 2;
    z.a = std::max<int64_t>(z.a, (int64_t)hw * hw * b.first);
    z.a = std::max<int64_t>(z.a, (int64_t)ho * ho * b.second);
    z.c = std::max<int64_t>(z.c, (int64_t)(hw + 1) * (hw + 1) * b.first);
    z.sd = std::max<int64_t>(z.sd, (int64_t)ho * ho * b.first);
    z.f = std::max<int64_t>(z.f, (int64_t)ho * ho * b.second);
    hw = ho;
  }
  return z;
}

static constexpr int64_t kSg2TailFloats = 2 * 16 * 512 + 512;   // last block output, final conv output, linear output (fp32)

static int64_t sg2_h16_bytes_per_sample(const sdg_ctx* c) {
  const Sg2Sizes z = sg2_h16_sizes(c);
  return 2 * (2 * z.a + z.c + z.sd) + 4 * kSg2TailFloats;
}

static int sg2_ensure_h16(sdg_ctx* c, int64_t chunk) {
  const Sg2Sizes z = sg2_h16_sizes(c);
  int rc;
  if ((rc = c->buf[0].ensure((size_t)chunk * z.a * 2))) return rc;
  if ((rc = c->buf[1].ensure((size_t)chunk * z.a * 2))) return rc;
  if ((rc = c->buf[2].ensure((size_t)chunk * z.c * 2))) return rc;
  if ((rc = c->buf[3].ensure((size_t)chunk * z.sd * 2))) return rc;
  if ((rc = c->buf[5].ensure((size_t)chunk * kSg2TailFloats * 4))) return rc;
  return 0;
}

static int forward_stylegan2_h16(sdg_ctx* c, const void* x, int layout, int64_t nb, float* logits, cudaStream_t s) {
  const int f16 = c->precision == SDG_PREC_FP16;
  const int S = c->size;
  h16* A = c->buf[0].as<h16>();
  h16* B = c->buf[1].as<h16>();
  h16* Cb = c->buf[2].as<h16>();
  h16* Sd = c->buf[3].as<h16>();
  float* tail = c->buf[5].as<float>();
  float* t_blk = tail;                               // [nb,16,512] last ResBlock output, fp32
  float* t_fc = t_blk + nb * 16 * 512;               // [nb,16,512] final conv output, fp32
  float* t_lin = t_fc + nb * 16 * 512;               // [nb,512]
  int rc, li = 0;
  {
    const ConvLayer& l = c->convs[li++];
    if ((rc = sg2_first_conv_h16(x, layout, l.w32.as<float>(), l.bias.as<float>(), A, nb, S, l.cout, f16, s))) return rc;
  }
  int hw = S;
  const size_t nblk = c->sg2_blocks.size();
  for (size_t bi = 0; bi < nblk; ++bi) {
    const auto& b = c->sg2_blocks[bi];
    const ConvLayer& c1 = c->convs[li++];
    const ConvLayer& c2 = c->convs[li++];
    li++;                                   // the skip conv's weights live in conv2's K extension
    const int ho = hw / 2;
    const bool last = bi + 1 == nblk;
    TcConv a1;
    a1.n = nb; a1.H = hw; a1.W = hw; a1.Cin = b.first; a1.Cout = b.first; a1.taps = 9;
    a1.in = A; a1.wb = c1.w16.as<h16>(); a1.bias = c1.bias.as<float>(); a1.act = 1; a1.out_raw = B;
    if (bi == 0 && (rc = prof_begin(c, s))) return rc;
    if ((rc = conv_tc(a1, f16, s))) return rc;
    if (bi == 0 && (rc = prof_end(c, s, 2.0 * (double)nb * hw * hw * b.first * 9.0 * b.first))) return rc;
    if ((rc = blur_h16(A, Sd, nb, hw, hw, b.first, 1, 2, f16, s))) return rc;
    if ((rc = blur_h16(B, Cb, nb, hw, hw, b.first, 2, 1, f16, s))) return rc;
    TcConv a2;
    a2.n = nb; a2.H = ho; a2.W = ho; a2.in_H = hw + 1; a2.in_W = hw + 1; a2.stride = 2; a2.no_pad = 1;
    a2.Cin = b.first; a2.Cout = b.second; a2.taps = 9;
    a2.in = Cb; a2.wb = c2.w16.as<h16>(); a2.bias = c2.bias.as<float>(); a2.act = 1;
    a2.sc_in = Sd; a2.sc_C = b.first; a2.sc_sep = 1; a2.out_scale = 0.70710678118654752f;
    if (last) a2.out_f32 = t_blk;          // fp32: the minibatch-stddev statistic and the split-precision tail read it
    else a2.out_raw = A;
    if ((rc = conv_tc(a2, f16, s))) return rc;
    hw = ho;
  }
  // tail: minibatch-stddev statistic (fp32), final conv (512 -> 512 on 4x4 + the stddev channel as a rank-1 term in the
  // epilogue), EqualLinear(8192, 512) as a plain GEMM, EqualLinear(512, 1) as a dot product
  float* sd = c->sg2_sd.as<float>();
  if ((rc = minibatch_stddev_fp32(t_blk, sd, nb, c->sg2_batch, 16, 512, s))) return rc;
  // split-precision tail: fp32 activations enter the GEMMs as [hi | lo | hi] 16-bit triples (B and Cb are free here)
  if ((rc = split3_rows_h16(t_blk, B, nb * 16, 512, f16, s))) return rc;
  const ConvLayer& fc = c->convs[li++];
  TcConv af;
  af.n = nb; af.H = 4; af.W = 4; af.Cin = 3 * 512; af.Cout = 512; af.taps = 9;
  af.in = B; af.wb = fc.w16.as<h16>(); af.bias = fc.bias.as<float>(); af.act = 1;
  af.sd = sd; af.sd_w = fc.w3.as<float>(); af.out_f32 = t_fc;
  if ((rc = conv_tc(af, f16, s))) return rc;
  if ((rc = split3_rows_h16(t_fc, Cb, nb * 16, 512, f16, s))) return rc;
  const ConvLayer& l0 = c->convs[li++];
  TcConv al;
  al.gemm = 1; al.n = 1; al.H = 1; al.W = (int)nb; al.Cin = 3 * 8192; al.Cout = 512; al.taps = 1;
  al.in = Cb; al.wb = l0.w16.as<h16>(); al.bias = l0.bias.as<float>(); al.act = 1; al.out_f32 = t_lin;
  if ((rc = conv_tc(al, f16, s))) return rc;
  return head_dot_fp32(t_lin, c->head_w.as<float>(), c->head_b.as<float>(), logits, nb, 512, s);
}

extern "C" int sdg_set_conv_pair(int on) {
  conv_tc_set_pair(on < 0 ? 0 : (on > 2 ? 1 : on));
  return 0;
}

extern "C" int sdg_ctx_profile(sdg_ctx* c, int enable) {
  SDG_REQUIRE(c, SDG_E_INVALID, "sdg_ctx_profile: null ctx");
  c->profile = enable != 0;
  c->prof_used = 0;
  c->prof_flops = 0.0;
  return 0;
}

extern "C" int sdg_ctx_profile_read(sdg_ctx* c, double* ms_total_host, int64_t* launches_host, double* flops_host) {
  SDG_REQUIRE(c && ms_total_host && launches_host && flops_host, SDG_E_INVALID, "sdg_ctx_profile_read: null pointer");
  double ms = 0.0;
  for (size_t i = 0; i < c->prof_used; ++i) {
    SDG_CUDA(cudaEventSynchronize(c->prof_events[i].second));
    float t = 0.f;
    SDG_CUDA(cudaEventElapsedTime(&t, c->prof_events[i].first, c->prof_events[i].second));
    ms += t;
  }
  *ms_total_host = ms;
  *launches_host = (int64_t)c->prof_used;
  *flops_host = c->prof_flops;
  c->prof_used = 0;
  c->prof_flops = 0.0;
  return 0;
}
