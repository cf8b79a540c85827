/* sdg.h -- C ABI of the B200-native per-sample diagnosis path of Self-Diagnosing GAN.
 *
 * The reference (grayhong/self-diagnosing-gan) has no FFI layer of its own: its boundary is the
 * Python surface of `diagan-pkg`.  Each entry point below names the
 * reference interface whose arithmetic it replaces (paths relative to the reference root); the
 * ctypes binding a maintainer adds on the reference side is shown in INTEGRATION.md and lives in
 * `self-diagnosing-gan_b200/diagan_b200/_lib.py`.
 *
 * Conventions
 *   - every pointer is a CALLER-OWNED DEVICE pointer unless the name ends in `_host`;
 *     the library never frees or retains them past the call (packed weights and scratch live in
 *     the opaque `sdg_ctx`);
 *   - `stream` is a `cudaStream_t` passed as `void*` (torch's current stream); all calls are
 *     asynchronous with respect to the host unless stated otherwise;
 *   - return value: 0 = ok, < 0 = invalid argument / unsupported shape (SDG_E_*), > 0 = cudaError_t.
 *     Nothing throws across the ABI.  `sdg_last_error()` returns a thread-local message;
 *   - one ctx per device/rank; calls on a ctx are not re-entrant;
 *   - there is NO CPU fallback anywhere behind this ABI.
 */
#ifndef SDG_H_
#define SDG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDG_ABI_VERSION 1

#define SDG_E_INVALID      (-1)   /* bad argument */
#define SDG_E_UNSUPPORTED  (-2)   /* shape / mode not implemented */
#define SDG_E_STATE        (-3)   /* call order (e.g. forward before load) */
#define SDG_E_DEVICE       (-4)   /* not an sm_100 device / driver entry point missing */

/* discriminator architectures */
#define SDG_ARCH_DCGAN32   1      /* diagan/models/mnist.py:155-223, eval mode, 3x32x32 */
#define SDG_ARCH_STYLEGAN2 2      /* diagan/models/stylegan2.py:619-677 (= stylegan2/model.py:603-660), any power-of-two size */
#define SDG_ARCH_SNGAN32   32     /* torch_mimicry SNGANDiscriminator32 (predefined_models.py:14,38,50) */
#define SDG_ARCH_SNGAN64   64     /* torch_mimicry SNGANDiscriminator64 (predefined_models.py:76,88) */

/* arithmetic of the conv stack */
#define SDG_PREC_FP32      0      /* CUDA-core fp32 (IEEE, no TF32): the 1e-5 parity mode */
#define SDG_PREC_BF16      1      /* tcgen05 kind::f16, bf16 operands, fp32 TMEM accumulators */
#define SDG_PREC_FP16      2      /* tcgen05 kind::f16, fp16 operands, fp32 TMEM accumulators: same speed as bf16,
                                     8x smaller rounding error -- the throughput mode.  Measured against the float64
                                     oracle: <= 1e-3 of the logit SCALE on every tested network; relative
                                     to an individual near-zero logit it exceeds 1e-3 (1.26e-2 on a random-init SNGAN-32
                                     whose logits cancel to ~-0.2, where cuDNN TF32 -- the reference's own GPU arithmetic --
                                     measures 1.42e-2).  fp16 has a finite RANGE (65504): see sdg_ctx_set_range_flag */

/* fp16 range guard: bits OR-ed into the caller's flag by any kernel that rounds to an fp16 operand */
#define SDG_RANGE_ACT      1      /* an activation left the fp16 range (|v| > 65504, or inf/NaN) when stored as an operand */
#define SDG_RANGE_WEIGHT   2      /* a packed weight did */

/* input layouts of sdg_d_forward */
#define SDG_LAYOUT_U8_NHWC   0    /* uint8 [n,H,W,3]; normalised (x/255-.5)/.5 on load (transform.py:3-11) */
#define SDG_LAYOUT_F32_NCHW  1    /* float32 [n,3,H,W] already in [-1,1] (what netD(x) receives, trainer.py:150) */

#if defined(__GNUC__)
#define SDG_API __attribute__((visibility("default")))
#else
#define SDG_API
#endif

typedef struct sdg_ctx sdg_ctx;

SDG_API const char* sdg_last_error(void);
SDG_API int sdg_abi_version(void);

/* per-device context: packed weights, TMA descriptors, activation scratch */
SDG_API int sdg_ctx_create(int device, sdg_ctx** out);
SDG_API int sdg_ctx_destroy(sdg_ctx* ctx);
/* fp16 range guard (no reference counterpart: the reference computes in fp32/TF32, whose range cannot be exceeded).
 * `device_flag` is a caller-owned, caller-zeroed DEVICE int32 (or NULL to detach).  With SDG_PREC_FP16 every kernel behind
 * sdg_*_load / sdg_d_forward that rounds a value to an fp16 operand ORs SDG_RANGE_* into it when the value is outside the
 * fp16 range, so that one overflowing pixel can never silently poison a logit and the running statistics after it: the
 * caller reads the flag after the pass and re-runs it with SDG_PREC_BF16 (fp32 range) -- diagan_b200 does so automatically
 * (LogitRecorder, range_check).  The flag is sticky; the library never clears it.  bf16 / fp32 modes never set it. */
SDG_API int sdg_ctx_set_range_flag(sdg_ctx* ctx, int32_t* device_flag);
/* upper bound on samples processed per internal sweep (bounds scratch memory); 0 = default */
SDG_API int sdg_ctx_set_chunk(sdg_ctx* ctx, int64_t samples_per_chunk);

/* ---- discriminator weights ------------------------------------------------------------------
 * Replaces torch_mimicry SNConv2d/SNLinear.sn_weights as called in eval mode under
 * trainer.py:145-150: one power iteration from the stored `sn_u` (buffers NOT updated), sigma =
 * u'.W.v, weight used = W / sigma.  sigma is identical for every batch of a recording pass, so it is
 * computed ONCE here and folded into the packed weights.
 * W/b/u: host arrays of `n_layers` DEVICE pointers in forward order
 *   SNGAN32: b1.c1 b1.c2 b1.c_sc b2.c1 b2.c2 b2.c_sc b3.c1 b3.c2 b4.c1 b4.c2 l5        (11)
 *   SNGAN64: b1.c1 b1.c2 b1.c_sc then (c1 c2 c_sc) for b2..b5, l6                       (16)
 * weights are torch layout [Cout,Cin,k,k] fp32 contiguous; u is [Cout].
 * inplace_relu != 0 reproduces mimicry's nn.ReLU(True) aliasing in DBlock (shortcut sees relu(x)). */
SDG_API int sdg_sngan_load(sdg_ctx* ctx, int arch, int n_layers,
                   const float* const* W_host, const float* const* b_host, const float* const* u_host,
                   int precision, int inplace_relu, void* stream);
/* copies the per-layer sigma computed by the last sdg_sngan_load to a device array [n_layers] */
SDG_API int sdg_sngan_sigmas(sdg_ctx* ctx, float* sigma_out, void* stream);

/* Replaces MNIST_DCGAN_Discriminator.forward in eval mode (mnist.py:213-223): BatchNorm running
 * statistics are folded into the conv weights/bias, Dropout is the identity.
 * conv_w: 6 device pointers [Cout,Cin,3,3]; bn_{gamma,beta,mean,var}: 5 device pointers each
 * (convs 2..6); fc_w [1,8192] in NCHW flatten order, fc_b [1]. */
SDG_API int sdg_dcgan_load(sdg_ctx* ctx, const float* const* conv_w_host,
                   const float* const* bn_gamma_host, const float* const* bn_beta_host,
                   const float* const* bn_mean_host, const float* const* bn_var_host,
                   const float* fc_w, const float* fc_b, int precision, void* stream);

/* Replaces StyleGANDiscriminator.forward (diagan/models/stylegan2.py:659-677; twin stylegan2/model.py:642-660) for an
 * input of `size` x `size`, channel_multiplier 2, blur kernel [1,3,3,1]: equalised-lr scales folded into the packed
 * weights, FusedLeakyReLU (op/fused_act.py), Blur = upfirdn2d (op/upfirdn2d.py) and minibatch-stddev reproduced.
 * tensors_host: device pointers in this order
 *   convs.0.0.weight, convs.0.1.bias, then per ResBlock i = 1..: conv1.0.weight, conv1.1.bias, conv2.1.weight, conv2.2.bias,
 *   skip.1.weight, then final_conv.0.weight, final_conv.1.bias, final_linear.0.weight, .0.bias, final_linear.1.weight, .1.bias
 * precision: SDG_PREC_FP32 = exact CUDA-core path; SDG_PREC_FP16 / SDG_PREC_BF16 = ResBlock convolutions on tcgen05
 * (16-bit NHWC activations, fp32 accumulation; the 3-channel first conv and the 4x4 tail stay fp32).
 * Minibatch-stddev couples the samples of a reference batch: sdg_ctx_set_batch gives the batch size B
 * the reference loader used (default 4); sdg_d_forward then needs n % B == 0 and treats samples [kB, (k+1)B) as one batch. */
SDG_API int sdg_stylegan2_load(sdg_ctx* ctx, int size, int n_tensors, const float* const* tensors_host, int precision,
                       void* stream);
SDG_API int sdg_ctx_set_batch(sdg_ctx* ctx, int batch);

/* ---- recording pass --------------------------------------------------------------------------
 * Replaces the loop body of LogTrainer._get_logit (trainer.py:148-154) and of stylegan2
 * get_logit (train_ffhq.py:135-141): logits_out[i] = netD(x_i) for i in [0,n), fp32.
 * The caller passes logits_out already offset to the first sample's dataset index. */
SDG_API int sdg_d_forward(sdg_ctx* ctx, const void* x, int layout, int64_t n, float* logits_out, void* stream);

/* One fused residual-block stage of the 16-bit tensor-core path on its own (what sdg_d_forward launches per
 * conv layer; exposed for kernel-level parity tests and the roofline measurement in bench.py).
 * Replaces, from torch-mimicry DBlock / DBlockOptimized (resblocks.py):
 *   v = F.conv2d(in, W/sigma, b, stride 1, padding ks/2)
 *       [+ F.conv2d(sc_in, Wsc/sigma)      the block's 1x1 shortcut conv: sc_C extra K columns of wb]
 *   [v = F.avg_pool2d(v, 2)                 pool != 0; pool == 2 evaluates conv + pool as the algebraically equal
 *                                           4x4 stride-2 conv (16/36 of the MACs): wb is then [Cout, 16*Cin + 4*sc_C] with
 *                                           wb[o][(a*4+b)*Cin + c] = 0.25 * sum_{ky in {a-1,a}, kx in {b-1,b}} W[o][c][ky][kx]
 *                                           and the shortcut block 0.25 * Wsc[o][c] repeated for its 4 taps]
 *   [v += sc_w3 . avg_pool2d(normalise(img), 2)   DBlockOptimized shortcut, img = the network input]
 *   [v += res_f32 (rectified if res_relu)   identity shortcut]
 *   out_relu = relu(v) 16-bit, out_raw = v 16-bit, out_f32 = v fp32        (each optional, >= 1 required)
 * in  [n,H,W,Cin] NHWC 16-bit (Cin % 64 == 0, H == W a power of two in 4..128; pooled: <= 64)
 * wb  [Cout, ks*ks*Cin + sc_C] 16-bit, K index = (ky*ks+kx)*Cin + c, then the shortcut columns
 * bias fp32 [Cout] (all biases summed) or NULL; Cout % 64 == 0, <= 1024; ks in {1,3}
 * precision: SDG_PREC_BF16 or SDG_PREC_FP16 = the 16-bit element type of in / wb / sc_in / out_relu / out_raw. */
SDG_API int sdg_conv2d_h16(const void* in, const void* wb, const float* bias, int64_t n, int H, int W, int Cin, int Cout,
                   int ks, const void* sc_in, int sc_C, int pool, const float* res_f32, int res_relu,
                   const void* img, int img_layout, const float* sc_w3, void* out_relu, void* out_raw,
                   float* out_f32, int precision, void* stream);
/* StyleGAN2 ConvLayer stages of the tensor-core path on their own (what sdg_d_forward launches per ResBlock conv; exposed
 * for kernel-level parity tests).  Replaces ConvLayer / ResBlock.forward pieces (stylegan2.py:553-616):
 *   v = F.conv2d(in, W * scale, stride = stride, padding = pad ? ks/2 : 0)          ks in {1,3}, stride in {1,2}
 *   [v = leaky_relu(v + bias, 0.2) * sqrt(2)     act != 0: FusedLeakyReLU (op/fused_act.py:104-116); else v += bias]
 *   [v += F.conv2d(skip_in, Wskip)               skip_in [n,Hout,Wout,skip_C] 16-bit at OUTPUT resolution (the blurred,
 *                                                decimated block input); Wskip = the last skip_C columns of wb; runs as extra
 *                                                K iterations into a second TMEM accumulator, added after the activation]
 *   [v += res_f32]  v *= out_scale               ResBlock: (out + skip) / sqrt(2)
 * in [n,in_H,in_W,Cin] NHWC 16-bit; outputs [n,Hout,Wout,Cout] (Hout == Wout a power of two in 4..512); wb
 * [Cout, ks*ks*Cin + skip_C] 16-bit with the equalised-lr scales folded in; out_raw 16-bit and/or out_f32. */
SDG_API int sdg_conv2d_sg2_h16(const void* in, const void* wb, const float* bias, int64_t n, int Hout, int Wout, int in_H,
                       int in_W, int Cin, int Cout, int ks, int stride, int pad, int act, const void* skip_in, int skip_C,
                       const float* res_f32, float out_scale, void* out_raw, float* out_f32, int precision, void* stream);
/* Blur (stylegan2.py:75-90 = upfirdn2d with outer([1,3,3,1])/64, zero padding `pad` on every side) on 16-bit NHWC;
 * out extent (H + 2*pad - 4) / stride + 1: stride 2 evaluates only the outputs a following stride-2 1x1 conv reads. */
SDG_API int sdg_blur_h16(const void* in, void* out, int64_t n, int H, int W, int C, int pad, int stride, int precision,
                 void* stream);
/* Kernel selection of the tensor-core convolution (process-wide; for tests and A/B timing):
 *   1 (default) every kernel where it applies: role-swapped (M = 128 channels x N = 256 pixels) for Cout = 128, streamed
 *               CTA pairs (tcgen05.mma.cta_group::2, N = 256) for Cout % 256 == 0, single-CTA otherwise;
 *   2           CTA-pair kernels (resident / streamed weights) but no role swap;
 *   0           the single-CTA pixel-major kernel only. */
SDG_API int sdg_set_conv_pair(int on);
/* First conv of the SNGAN discriminators straight from the dataset bytes: out = relu(conv3x3(normalise(x)) + b).
 * Replaces transform.py:3-11 + DBlockOptimized.c1 + ReLU.  x: uint8 [n,S,S,3] or fp32 [n,3,S,S] (layout);
 * wb 16-bit [Cout][64] with K index (ky*3+kx)*3 + c (27 real columns, rest zero); S in {32,64}; Cout in {64,128}. */
SDG_API int sdg_first_conv_h16(const void* x, int layout, const void* wb, const float* bias, void* out, int64_t n, int S,
                       int Cout, int precision, void* stream);

/* SNGAN-32 block 1 (torch-mimicry DBlockOptimized(3, 128): c1 -> ReLU -> c2 -> avg_pool2d, + c_sc(avg_pool2d(x)); call site
 * diagan/trainer/trainer.py:150 through SNGANDiscriminator32.forward) as ONE launch: relu(c1(x)) is built and consumed in shared
 * memory (conv_b1fused.cu).  out_relu[n,16,16,128] = relu(block1(x)) 16-bit, what sdg_first_conv_h16 followed by
 * sdg_conv2d_h16(pool = 2, img) writes.  x: uint8 [n,32,32,3]; w1 [128][64] as for sdg_first_conv_h16; w2 [128][2048] in the
 * 4x4 stride-2 form (K index (a*4+b)*128 + c); bias2 [128] = c2 + shortcut bias; sc_w3 [128][3] fp32;
 * dbg_t: optional [n,32,32,128] copy of the intermediate tensor (tests), or null. */
SDG_API int sdg_sngan32_block1_fused_h16(const void* x, const void* w1, const float* b1, const void* w2, const float* bias2,
                                 const float* sc_w3, void* out_relu, void* dbg_t, int64_t n, int precision, void* stream);

/* SNGAN-64 block 1 (torch-mimicry DBlockOptimized(3, 64) of SNGANDiscriminator64; call sites predefined_models.py:76,88) as ONE
 * launch, per 32 x 32 quadrant of the image (halo recomputed from the neighbouring quadrants' pixels).  x: uint8 [n,64,64,3];
 * w1 [64][64]; w2 [64][1024] in the 4x4 stride-2 form (K index (a*4+b)*64 + c); bias2 [64]; sc_w3 [64][3];
 * out_relu [n,32,32,64]; dbg_t: optional [n,64,64,64] copy of relu(c1(x)) (tests), or null. */
SDG_API int sdg_sngan64_block1_fused_h16(const void* x, const void* w1, const float* b1, const void* w2, const float* bias2,
                                 const float* sc_w3, void* out_relu, void* dbg_t, int64_t n, int precision, void* stream);

/* ---- dataset transform (the input pipeline of the pass) ----------------------
 * Replaces transforms.Resize(size) + transforms.CenterCrop(size) of datasets/transform.py:3-41 as applied to every item of
 * every recording pass by the reference's DataLoader workers: one pass over the raw uint8 dataset in [n,H,W,C] (C = 3 or 1)
 * -> out [n,size,size,C] uint8, BIT-EXACT with Pillow's 8-bit bilinear ImagingResample (horizontal pass first, 22-bit
 * fixed point) and torchvision's Resize(int) / CenterCrop geometry.  ToTensor + Normalize(0.5, 0.5) are not materialised:
 * sdg_d_forward normalises uint8 input on load.  in/out are device pointers. */
SDG_API int sdg_resize_center_crop_u8(const uint8_t* in, int64_t n, int H, int W, int C, int size, uint8_t* out, void* stream);

/* ---- running per-sample statistics (new: the reference keeps every snapshot, trainer.py:337-338) --
 * Welford update with snapshot number t (0-based) plus last value and sum |x_t - x_{t-1}|:
 * after T updates  ldrm = mean, ldrv = m2/(T-1), ldr = last, ldrd = sad/(T-1)  (plot.py:243-246). */
SDG_API int sdg_stats_update(const float* snapshot, double* mean, double* m2, double* last, double* sad,
                     int64_t n, int64_t t, void* stream);

/* ---- scoring ---------------------------------------------------------------------------------
 * Replaces calculate_scores (plot.py:220-249).
 * sdg_window_moments_{f32,f64}: snaps [T, ld] row-major (row t = snapshot t of the selected window);
 * two-pass mean / ddof-1 variance in NumPy's evaluation order (bit-exact vs np.mean/np.var/np.std),
 * ldrd = mean_t |x_{t+1}-x_t|, ldr = last row.  Any output pointer may be NULL. */
SDG_API int sdg_window_moments_f32(const float* snaps, int64_t T, int64_t n, int64_t ld,
                           double* mean, double* var, double* ldrd, double* ldr, void* stream);
SDG_API int sdg_window_moments_f64(const double* snaps, int64_t T, int64_t n, int64_t ld,
                           double* mean, double* var, double* ldrd, double* ldr, void* stream);
/* phase 1: score[j][i] = max(mean[i] + conf[j]*sqrt(var[i]), floor), mins[j] = min_i score[j][i]
 * (clip_min, plot.py:230-231).  conf_host: host array of n_conf multipliers; score [n_conf, n];
 * mins [n_conf] device.  var_is_m2_over: if > 0, `var` holds Welford m2 and is divided by it. */
SDG_API int sdg_score_floor_min(const double* mean, const double* var, int64_t n, const double* conf_host,
                        int n_conf, double floor, double var_is_m2_over, double* score, double* mins,
                        void* stream);
/* phase 2: score[j][i] = max(min(score[j][i], mins[j]*ratio), eps)  (clip_max_ratio plot.py:226-228,
 * then the sampler floor of train_mimicry_phase2.py:23; pass eps = 0 to skip the floor).
 * Between the phases a sharded caller all-reduces `mins` with MIN. */
SDG_API int sdg_score_clip(double* score, int64_t n, int n_conf, const double* mins, double ratio, double eps,
                   void* stream);
/* One-collective form for sample-index shards (the reference's precedent, stylegan2/train_ffhq.py:128-143,
 * issues two all-gathers per batch of 4): every rank runs phase 1 on its shard with score = payload, mins = payload +
 * shard_size, all-gathers the [shard_size + 1] doubles, and this call clips the gathered [world][shard_size + 1] buffer
 * against the GLOBAL minimum (min over the world trailing slots) into the dense vector out [n]:
 * out[i] = max(min(v_i, gmin * ratio), eps).  A rank with an empty shard contributes +inf as its minimum. */
SDG_API int sdg_score_clip_gathered(const double* gathered, int world, int64_t shard_size, int64_t n, double ratio,
                            double eps, double* out, void* stream);

/* ---- top-index selection ---------------------------------------------------------------------
 * Replaces np.argsort(w)[-k:] / [:k] (eval_gan_drs_with_index.py:97-99, plot.py:100-101) with the
 * tie-break of a stable sort: idx_out[0..k) in ascending (score, index) order, equal to
 * np.argsort(w, kind="stable")[-k:] (largest != 0) or [:k].  k <= 4096.
 * workspace: device scratch of at least sdg_topk_workspace_bytes(n) bytes. */
SDG_API size_t sdg_topk_workspace_bytes(int64_t n);
SDG_API int sdg_topk_indices(const double* score, int64_t n, int k, int largest, int64_t* idx_out,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---- Discriminator Rejection Sampling ----------------------------------------------------------
 * Replaces DRS.init_drs / sub_rejection_sampler (models/drs.py:31-57, trainer/evaluate.py:45-68).
 * running_max: device float, initialised by the caller to -100000.
 * sdg_drs_update_max: running_max = max(running_max, max_i ldr[i])      (burn-in, drs.py:31-36)
 * sdg_drs_accept: updates running_max, F = l - log(1 - exp(l - eps)) with l = ldr - max (fp32),
 *   gamma = `gamma` if use_gamma else the `percentile`-th linear-interpolation percentile of F,
 *   p = sigmoid(F - gamma), accept[i] = p[i] > psi[i]; accepted indices are compacted in order into
 *   idx_out and their number written to count_out.  n <= 2048.  Any of p_out/accept_out/idx_out may
 *   be NULL. */
SDG_API int sdg_drs_update_max(const float* ldr, int n, float* running_max, void* stream);
SDG_API int sdg_drs_accept(const float* ldr, int n, float* running_max, float eps, float percentile,
                   int use_gamma, float gamma, const double* psi,
                   float* p_out, uint8_t* accept_out, int32_t* idx_out, int32_t* count_out, void* stream);

/* ---- introspection for bench.py ("gpu_launches") ---------------------------------------------- */
SDG_API int64_t sdg_launch_count(void);       /* kernels launched by this library since load / last reset */
SDG_API void sdg_launch_count_reset(void);

/* Event timing of the dominant kernel of the bf16 SNGAN forward (block1.c2, the 3x3 128->128 conv at
 * full resolution) on the launching stream, for the roofline line of bench.py.  _read synchronises on
 * the recorded events, returns total milliseconds, number of launches and the FLOPs those launches EXECUTED
 * (2*M*N*K of the GEMM actually run: with the 4x4 stride-2 form of conv3x3 + avg_pool2d that is 16/36 of the
 * reference formulation's count) since the last read, and clears the record. */
SDG_API int sdg_ctx_profile(sdg_ctx* ctx, int enable);
SDG_API int sdg_ctx_profile_read(sdg_ctx* ctx, double* ms_total_host, int64_t* launches_host, double* flops_host);

#ifdef __cplusplus
}
#endif
#endif /* SDG_H_ */
