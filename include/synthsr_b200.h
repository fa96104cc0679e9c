/* libsynthsr_b200 -- C ABI of the B200-native SynthSR training hot path.
 *
 * The reference (BBillot/SynthSR) has no FFI: its hot path is a Keras graph.  This header is the drop-in
 * boundary a maintainer binds from Python (ctypes); every entry point names the reference
 * code it replaces (file:line relative to the reference repository).
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no torch types.  All data pointers are DEVICE pointers unless noted.
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), allocates nothing and keeps no state
 *    other than cached TMA descriptors; scratch buffers are provided by the caller.
 *  - return value: 0 on success, <0 on error (-1 bad argument, -2 CUDA error, -3 unsupported); the message is in
 *    ssr_last_error() (thread local).
 *  - volumes are [B][d0][d1][d2][C] float32 (channels contiguous), labels int32 [B][d0][d1][d2]; (d0,d1,d2) are the
 *    reference's three spatial axes in order, d2 contiguous.
 *  - *_stride/*_off pairs address a channel slice of a wider channels-last tensor (element (v,c) at
 *    v*stride + off + c); stride <= 0 means "dense".
 */
#ifndef SYNTHSR_B200_H_
#define SYNTHSR_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- library ---------------------------------- */
const char* ssr_last_error(void);
unsigned long long ssr_launch_count(void);     /* kernels launched by this library since load */
int ssr_abi_version(void);
int ssr_device_arch(void);                     /* 100 on B200 */

/* ---------------------------------------------------------------- generator -------------------------------- */
/* nrn_layers.Resize (ext/neuron/layers.py:361-394 -> ext/neuron/utils.py:127-154), linear or nearest, C<=3. */
int ssr_resize(const float* src, float* dst, int B, int s0, int s1, int s2, int d0, int d1, int d2, int C,
               int nearest, int dst_stride, int dst_off, void* stream);

/* nrn_layers.VecInt scaling-and-squaring (ext/neuron/layers.py:241-272 -> utils.integrate_vec :351-369).
 * vec [B][n0][n1][n2][3] in/out, tmp same size. */
int ssr_svf_integrate(float* vec, float* tmp, int B, int n0, int n1, int n2, int nb_steps, void* stream);

/* Fused PadAroundCentre + Resize(field) + SpatialTransformer('nearest') + RandomCrop + RandomFlip(+L/R swap):
 * ext/lab2im/layers.py:196-211, 252-270, 391-427, 1754; ext/neuron/utils.py:222-286, 112-122.  Bit exact.
 * labels [B][n0-2p0][n1-2p1][n2-2p2]; aff [B][4][4] or NULL; field_half [B][h0][h1][h2][3] (integrated SVF) or
 * NULL (then h*=0); crop_idx [B][3] or NULL; flip [B] bytes or NULL; swap_lut [lut_len] or NULL;
 * out [B][c0][c1][c2]. */
int ssr_deform_labels_nearest(const int* labels, int* out, const float* aff, const float* field_half, int B, int n0,
                              int n1, int n2, int p0, int p1, int p2, int h0, int h1, int h2, const int* crop_idx,
                              int c0, int c1, int c2, const unsigned char* flip, const int* swap_lut, int lut_len,
                              void* stream);

/* Same transform with SpatialTransformer('linear') on a single-channel float image (real-image target,
 * registration-error simulation: SynthSR/labels_to_image_model.py:128-134, 202-208, 231-238). */
int ssr_warp_linear(const float* image, float* out, const float* aff, const float* field_half, int B, int n0, int n1,
                    int n2, int p0, int p1, int p2, int h0, int h1, int h2, const int* crop_idx, int c0, int c1, int c2,
                    const unsigned char* flip, void* stream);

/* tf.random.normal replacement: Philox4x32-10 + Box-Muller, n standard normals. */
int ssr_philox_normal(float* out, long long n, unsigned long long seed, unsigned long long stream_id, void* stream);

/* Fused SampleConditionalGMM + BiasFieldCorruption + clip + per-item min/max for ONE synthetic channel:
 * ext/lab2im/layers.py:480-498, 1067-1097, 1214-1215, 1230-1231.
 * lut_mean/lut_std [B][lut_len] (scatter of means/stds at generation_labels); noise [B][V] standard normals or NULL
 * (then generated on the fly from seed/stream_id); bias_small [B][b0][b1][b2] (N(0,std) draws) or NULL;
 * minmax [B][2] uint32 (order-preserving encoding, consumed by ssr_blur3d). */
int ssr_gmm_bias_minmax(const int* labels, const float* lut_mean, const float* lut_std, int lut_len,
                        const float* noise, unsigned long long seed, unsigned long long stream_id,
                        const float* bias_small, int b0, int b1, int b2, int apply_bias, float clip_max, float* out,
                        unsigned int* minmax, int B, int n0, int n1, int n2, void* stream);

/* per-item min/max of a float volume (IntensityAugmentation(normalise=True) on a real image). */
int ssr_minmax(const float* x, unsigned int* minmax, int B, long long nvox, void* stream);

/* GaussianBlur / tf.nn.conv3d(...,'SAME') with a dense k0 x k1 x k2 kernel (ext/lab2im/layers.py:732-767), with
 * the min-max normalisation and gamma augmentation of IntensityAugmentation (layers.py:1235-1242) optionally fused
 * on the loads (minmax / gamma_exp[B] non-NULL).  Single channel in, single channel (slice) out. */
int ssr_blur3d(const float* src, float* dst, const float* kern, int k0, int k1, int k2, const unsigned int* minmax,
               const float* gamma_exp, int B, int n0, int n1, int n2, int src_stride, int src_off, int dst_stride,
               int dst_off, void* stream);

/* MimicAcquisition (ext/lab2im/layers.py:835-999; randomise_res branch of labels_to_image_model.py:215-220): nearest
 * resampling to a per-example acquisition grid + linear resampling to (o0,o1,o2), fused; params [B][9] = down_zoom[3] |
 * up_zoom[3] | resolution[3]; dist (optional) = distance in mm to the nearest acquired voxel. */
int ssr_mimic_acquisition(const float* src, float* dst, float* dist, const float* params, int B, int n0, int n1,
                          int n2, int o0, int o1, int o2, int dst_stride, int dst_off, int dist_stride, int dist_off,
                          void* stream);
int ssr_copy_strided(const float* src, float* dst, long long n, int src_stride, int src_off, int dst_stride,
                     int dst_off, void* stream);

/* reliability map (ext/lab2im/edit_tensors.py:313-333): outer product of per-axis factors (double, device) or 1. */
int ssr_fill_outer3(float* dst, const double* f0, const double* f1, const double* f2, int B, int n0, int n1, int n2,
                    int dst_stride, int dst_off, void* stream);

/* ---------------------------------------------------------------- U-Net: exact fp32 convolutions ------------ */
/* KL.Conv3D(Cout, k, padding='same', activation) on the logical concat [x1, x2] (ext/neuron/models.py:316,444). */
int ssr_conv3d_fwd_ref(const float* x1, int C1, const float* x2, int C2, const float* w, const float* bias, float* y,
                       int B, int d0, int d1, int d2, int Cout, int k, int act, void* stream);
int ssr_conv3d_dgrad_ref(const float* dy, const float* w, float* wd_scratch, float* dx, int B, int d0, int d1, int d2,
                         int Cin, int Cout, int k, void* stream);
int ssr_conv3d_wgrad_ref(const float* x1, int C1, const float* x2, int C2, const float* dy, float* dw, float* db, int B,
                         int d0, int d1, int d2, int Cout, int k, void* stream);

/* ---------------------------------------------------------------- U-Net: tcgen05 (TF32) convolutions -------- */
/* Same contract as the *_ref functions for k == 3, computed on the 5th-generation tensor cores
 * (tcgen05.mma.kind::tf32, TMA-staged shared-memory tiles, TMEM accumulators).  See conv_tc.cu. */
int ssr_conv3d_pack_weights(const float* w, float* wp, int Cin1, int Cin2, int Cout, int mode, void* stream);
long long ssr_conv3d_packed_size(int Cin1, int Cin2, int Cout, int mode);
/* jobs: DEVICE int64 array, njobs x {w pointer, wp pointer, Cin1, Cin2, Cout, mode}: packs every layer in one launch */
int ssr_conv3d_pack_weights_batch(const long long* jobs, int njobs, void* stream);
int ssr_conv3d_fwd_tc(const float* x1, int C1, const float* x2, int C2, const float* wp, const float* bias, float* y,
                      int B, int d0, int d1, int d2, int Cout, int act, void* stream);
/* forward that also accumulates the BatchNorm sums of its output in the epilogue (sums: 2*Cout doubles, zeroed here: sum |
 * sum of squares; finish with ssr_bn_finalize) -- KL.BatchNormalization statistics (models.py:349-351, 475-477) */
int ssr_conv3d_fwd_tc_stats(const float* x1, int C1, const float* x2, int C2, const float* wp, const float* bias, float* y,
                            double* sums, int B, int d0, int d1, int d2, int Cout, int act, void* stream);
/* data gradient (wp: pack mode 1) fused with the ELU backward of the layer below: dx = conv(dy, wp) * elu'(h),
 * dbias[c] += sum_v dx[v][c]; h = that layer's forward output, Cout = channels of dx */
int ssr_conv3d_dgrad_tc_elu(const float* dy, int C, const float* wp, const float* h, float* dx, float* dbias, int B, int d0,
                            int d1, int d2, int Cout, void* stream);
/* Small deep layers run split-K (the chunk list of K cut into parts that add their partial sums into y with red.global.add,
 * bias + activation in a follow-up pass): ssr_conv3d_fwd_tc / _acc / _comp do this on their own.  The fused-epilogue entry
 * points (*_stats, *_elu) never split; this query tells a caller which factor the plain entry point would use (1 = none) so
 * that it can prefer the plain call + separate BatchNorm / ELU' passes there.  comp: 0, or the compensation level. */
int ssr_conv3d_fwd_tc_ksplit(int C1, int C2, int Cout, int B, int d0, int d1, int d2, int comp);
/* same, added to the partial result already in y before bias + activation */
int ssr_conv3d_fwd_tc_acc(const float* x1, int C1, const float* x2, int C2, const float* wp, const float* bias, float* y,
                          int B, int d0, int d1, int d2, int Cout, int act, void* stream);
/* Decoder levels (UpSampling3D -> concatenate -> Conv3D, ext/neuron/models.py:425-446) without the upsampled tensor: per
 * output parity class the 3x3x3 kernel over a nearest-upsampled input is an effective 2x2x2 kernel over the LOW-resolution
 * input (8 taps instead of 27).  ssr_conv3d_up_weights forms the effective kernels weff[8][27][Cup][Cout] (unused taps
 * zero) and a contiguous copy of the skip part; they are packed per parity class with ssr_conv3d_pack_weights (mode 0 for
 * the forward, mode 1 for the gradient), 8 packs back to back.
 *   fwd_tc_up:   y[B,2d0,2d1,2d2,Cout] = partial sums of the upsampled part (no bias / activation; add the skip part with
 *                ssr_conv3d_fwd_tc_acc / ssr_conv3d_fwd_tc_k2n_part(accumulate))
 *   dgrad_tc_up: dlow[B,d0,d1,d2,Cup] = gradient w.r.t. the low-resolution tensor (UpSampling3D backward included) */
int ssr_conv3d_up_weights(const float* w, int Cskip, int Cup, int Cout, float* wskip, float* weff, void* stream);
int ssr_conv3d_fwd_tc_up(const float* low, int Cup, const float* wp8, float* y, int B, int d0, int d1, int d2, int Cout,
                         void* stream);
int ssr_conv3d_dgrad_tc_up(const float* dy, int Cout_layer, const float* wp8, float* dlow, int B, int d0, int d1, int d2,
                           int Cup, void* stream);
/* ssr_conv3d_fwd_tc_up in the k2n layout for the last decoder level (Cout == 24, Cup <= 64): d2 parity and its two taps in
 * the MMA N dimension, the kernels of one (p0, p1) class resident per CTA; wpk: 4*8*96*32 floats from
 * ssr_conv3d_pack_up_k2n(weff of ssr_conv3d_up_weights) */
int ssr_conv3d_pack_up_k2n(const float* weff, float* wpk, int Cup, void* stream);
int ssr_conv3d_fwd_tc_up_k2n(const float* low, int Cup, const float* wpk, float* y, int B, int d0, int d1, int d2, int Cout,
                             void* stream);
/* Cin <= 32, Cout <= 32 (the full-resolution layers): the three d2 taps ride in the MMA N dimension; weights packed
 * with mode 2 (forward) / 3 (data gradient: x = dy, Cout = the layer's Cin).  Same result contract as ssr_conv3d_fwd_tc. */
int ssr_conv3d_fwd_tc_k2n(const float* x, int C, const float* wp, const float* bias, float* y, int B, int d0, int d1,
                          int d2, int Cout, int act, void* stream);
/* one <= 32-channel part [c0, c0 + C) of the input of a Cout <= 32 convolution (a concatenated input is the sum of its
 * parts): weights packed with mode 4 (Cin1 = total input channels, Cin2 = (c0 << 8) | C); accumulate = add to the partial
 * result in y, final = apply bias + activation. */
int ssr_conv3d_fwd_tc_k2n_part(const float* x, int Ctot, int c0, int C, const float* wp, const float* bias, float* y, int B,
                               int d0, int d1, int d2, int Cout, int act, int accumulate, int final, void* stream);
/* k2n forward that also accumulates the BatchNorm statistics of its output in the epilogue (sums: 2*Cout doubles,
 * zeroed here: sum | sum of squares; finish with ssr_bn_finalize).  Cout = 24 or 32.  Replaces the reduction pass of
 * KL.BatchNormalization over the full-resolution tensor (ext/neuron/models.py:349-351, 475-477). */
int ssr_conv3d_fwd_tc_k2n_stats(const float* x, int C, const float* wp, const float* bias, float* y, double* sums, int B,
                                int d0, int d1, int d2, int Cout, int act, void* stream);
/* k2n data gradient fused with the ELU backward of the convolution below (activation='elu', models.py:316,444):
 * dx = conv(dy, wp) * elu'(h), dbias[c] += sum_v dx[v][c]; h = that convolution's forward output.  Cout = 24 or 32. */
int ssr_conv3d_dgrad_tc_k2n_elu(const float* dy, int C, const float* wp, const float* h, float* dx, float* dbias, int B,
                                int d0, int d1, int d2, int Cout, void* stream);
/* Compensated forward ("3xTF32"): fp32-class accuracy of KL.Conv3D (fp32 in the reference, ext/neuron/models.py:316,444,
 * 481) on the TF32 tensor cores.  With x = x_hi + x_lo (x_hi = what the TMA's TFLOAT32 load makes of x, x_lo =
 * ssr_tf32_residual(x)) and w = w_hi + w_lo (pack mode 5: Cin1 = total input channels of the kernel, Cin2 = (first channel
 * << 12) | channels; pack mode 6: the lo part in the k2n layout of mode 2) the convolution is evaluated as ONE implicit
 * GEMM over K = [x | x_lo | x] against [w_hi | w_hi | w_lo] (level 3), or [x | x_lo] against [w_hi | w_hi] (level 2).
 * sums (or NULL): BatchNorm sums of the output, as ssr_conv3d_fwd_tc_stats; accumulate: add to the partial result in y. */
int ssr_tf32_residual(const float* x, float* lo, long long n, void* stream);
int ssr_conv3d_fwd_tc_comp(const float* x, const float* xlo, int C, const float* wp, const float* bias, float* y,
                           double* sums, int B, int d0, int d1, int d2, int Cout, int act, int accumulate, int level,
                           void* stream);
int ssr_conv3d_fwd_tc_up_comp(const float* low, const float* lowlo, int Cup, const float* wp8c, float* y, int B, int d0,
                              int d1, int d2, int Cout, int level, void* stream);
/* Hybrid scheme (level 4, the default of conv_impl='tc3'): the two correction terms run as ONE bf16 MMA chain -- x2 =
 * [bf16(x_lo) | bf16(x_hi)] (ssr_tf32_split_bf16: a 2C-channel bf16 tensor, same bytes as x) against [w_hi ; w_lo] in bf16
 * (pack mode 7 = TF32 hi chunks followed by the bf16 chunks; pack mode 8 = the bf16 chunk in the k2n layout).  The
 * corrections are ~2^-11 of the result, so bf16's 8-bit operands leave ~2^-20 relative -- and a bf16 MMA covers twice the
 * K of a TF32 one: 2 chains per convolution instead of 3.  x2 / lowlo of the *_comp entry points is then that bf16 tensor. */
int ssr_tf32_split_bf16(const float* x, void* x2, long long nvox, int C, void* stream);
int ssr_conv3d_fwd_tc_k2n_bf16(const void* x2, int C2, const float* wp, const float* bias, float* y, double* sums, int B,
                               int d0, int d1, int d2, int Cout, int act, void* stream);
/* bf16x3 scheme (level 5; generic and parity kernels): every term in bf16.  x = x1 + x2 (+ 2^-18), w = w1 + w2 (+ 2^-18)
 * with 8-bit pieces; the convolution is x1 w1 + x2 w1 + x1 w2 -- three bf16 K-chunks per 64 input channels = 1.5 TF32
 * chains (the hybrid scheme runs 2), the dropped terms are ~2^-17 of a product.  x2 / lowlo of the *_comp entry points is
 * ssr_bf16x3_split's [x1 | x2] (2C bf16 channels per voxel); the weights are pack mode 9 (w1 chunks, then w2 chunks; Cin1 /
 * Cin2 as in mode 5). */
int ssr_bf16x3_split(const float* x, void* x2, long long nvox, int C, void* stream);
/* producers that emit the next layer's x2 from their own epilogue (bit-identical to ssr_tf32_split_bf16 of their output,
 * without the extra pass): the first layer (KL.Conv3D with Cin <= 2) and the final k2n channel part */
int ssr_conv3d_first_fwd_split(const float* x, int C1, const float* w, const float* bias, float* y, void* y2, int B, int d0,
                               int d1, int d2, int Cout, int act, void* stream);
int ssr_conv3d_fwd_tc_k2n_part_split(const float* x, int Ctot, int c0, int C, const float* wp, const float* bias, float* y,
                                     void* y2, int B, int d0, int d1, int d2, int Cout, int act, int accumulate,
                                     void* stream);
/* last channel part of a k2n convolution + the BatchNorm sums of the finished output (sums zeroed here) */
int ssr_conv3d_fwd_tc_k2n_part_stats(const float* x, int Ctot, int c0, int C, const float* wp, const float* bias, float* y,
                                     double* sums, int B, int d0, int d1, int d2, int Cout, int act, int accumulate,
                                     void* stream);
int ssr_conv3d_wgrad_tc(const float* x1, int C1, const float* x2, int C2, const float* dy, float* dw, float* db,
                        float* scratch, long long scratch_bytes, int B, int d0, int d1, int d2, int Cout,
                        void* stream);
/* weight gradient w.r.t. the input channels [cin_off, cin_off + C) of a kernel dw (27, cin_total, Cout) */
int ssr_conv3d_wgrad_tc_part(const float* x, int C, const float* dy, float* dw, int cin_total, int cin_off, int B, int d0,
                             int d1, int d2, int Cout, void* stream);
/* upsampled part of a decoder convolution from the LOW-resolution tensor low [B,d0,d1,d2,Cup] and the full-resolution
 * dy [B,2d0,2d1,2d2,Cout]: gradients of the 8 effective kernels (scratch: 8*27*Cup*Cout floats, zeroed here) combined into
 * dw (27, cin_total, Cout) at input channels [cin_off, cin_off + Cup) */
int ssr_conv3d_wgrad_tc_up(const float* low, int Cup, const float* dy, float* dw, int cin_total, int cin_off, float* scratch,
                           int B, int d0, int d1, int d2, int Cout, void* stream);
long long ssr_conv3d_wgrad_scratch_bytes(int C1, int C2, int Cout, int B, int d0, int d1, int d2);
int ssr_tc_selftest(void* stream);
int ssr_tc_set_debug(long long* buf);   /* profiling: per-CTA clock64 phase stamps of conv3d_tc_kernel */
/* tcgen05.mma issue-cost microbenchmark (cycles per MMA per CTA); see profiles/ */
int ssr_tc_microbench(float* out, int nblocks, int N, int nacc, int chain, int iters, int kmajor, int commit_every,
                      int cycle_addr, void* stream);

/* shared-memory descriptor probe (which unaligned / overlapping operand views the tensor core reads consistently with
 * the TMA swizzle); a, b: [rows][32] fp32, d: [128][N]; p: HOST int[14], see conv_tc.cu */
int ssr_tc_desc_probe(const float* a, int rows_a, const float* b, int rows_b, float* d, int swz, const int* p,
                      void* stream);

/* ---------------------------------------------------------------- U-Net: other layers ----------------------- */
int ssr_channel_sum(const float* t, long long nvox, int C, float* out, void* stream);
/* KL.BatchNormalization(axis=-1), training mode (ext/neuron/models.py:349-351, 475-477).
 * stats [4*C] = mean | invstd | gamma*invstd | beta - mean*gamma*invstd ; sums_scratch 2*C doubles. */
int ssr_bn_stats(const float* x, long long nvox, int C, const float* gamma, const float* beta, float* moving_mean,
                 float* moving_var, float eps, float momentum, double* sums_scratch, float* stats, void* stream);
int ssr_bn_finalize(const double* sums, long long nvox, int C, const float* gamma, const float* beta, float* moving_mean,
                    float* moving_var, float eps, float momentum, float* stats, void* stream);
int ssr_bn_stats_inference(int C, const float* gamma, const float* beta, const float* moving_mean,
                           const float* moving_var, float eps, float* stats, void* stream);
/* mode 0: BN ; 1: BN + MaxPooling3D(2,'same') (models.py:354-356) ; 2: BN + UpSampling3D(2) (models.py:425-427). */
int ssr_bn_apply(const float* x, float* y, const float* stats, int B, int d0, int d1, int d2, int C, int mode,
                 int dst_stride, int dst_off, void* stream);
/* dbias (optional): += per-channel sum of the result (bias gradient of the convolution that produced x) */
int ssr_bn_bwd(const float* dy, const float* x, const float* stats, long long nvox, int C, const float* add,
               int add_stride, int add_off, int elu, float* dx, float* dgamma, float* dbeta, float* dbias,
               double* sums_scratch, void* stream);
/* encoder level: MaxPooling3D backward + BatchNormalization backward (+ skip gradient `add`, + ELU') in two passes, the
 * full-resolution unpooled gradient is never materialised (= ssr_maxpool_bwd followed by ssr_bn_bwd).  (d0,d1,d2) = shape
 * of x; dp = gradient w.r.t. the pooled output. */
int ssr_pool_bn_bwd(const float* dp, const float* x, const float* stats, int B, int d0, int d1, int d2, int C,
                    const float* add, int add_stride, int add_off, int elu, float* dx, float* dgamma, float* dbeta,
                    float* dbias, double* sums_scratch, void* stream);
/* ssr_bn_bwd with the two reductions already in sums2 = [sum dy | sum dy * xhat] (2*C doubles), e.g. from
 * ssr_head_loss_bnsums: no reduction pass over dy / x */
int ssr_bn_bwd_sums(const float* dy, const float* x, const float* stats, long long nvox, int C, const float* add,
                    int add_stride, int add_off, int elu, float* dx, float* dgamma, float* dbeta, float* dbias,
                    double* sums2, void* stream);
int ssr_maxpool_bwd(const float* dp, const float* x, const float* stats, int B, int d0, int d1, int d2, int C,
                    float* dy_full, void* stream);
int ssr_upsample_bwd(const float* du, int du_stride, int du_off, int B, int d0, int d1, int d2, int C, float* dlow,
                     void* stream);
int ssr_elu_bwd(const float* dh, int dh_stride, int dh_off, const float* h, const float* add, long long nvox, int C,
                float* da, float* dbias, void* stream);
/* unet_likelihood 1x1x1 conv (models.py:480-481) + metrics_model (SynthSR/metrics_model.py:53-104), fwd + bwd. */
/* feat_stats (optional): the 4*C BatchNorm stats of the layer that produced `feat`; the normalisation is then applied on
 * the fly (feat = raw * scale + shift) and the normalised tensor is never materialised. */
int ssr_head_loss(const float* feat, const float* feat_stats, const float* w, const float* bias, const float* image,
                  int image_channels,
                  const int* res_idx, const float* target, float* pred, float* dfeat, float* dw, float* db,
                  double* loss, float* gout_scratch, int B, int d0, int d1, int d2, int C, int L, int metric,
                  const int* crop_size, const int* crop_begin, int train, void* stream);
/* training-mode ssr_head_loss with the BatchNorm folded in (feat_stats required) that also returns the two reductions of
 * that BatchNorm's backward, sums2 = [sum_v dfeat | sum_v dfeat * xhat] (2*C doubles), obtained algebraically from the head
 * gradients (dfeat = g w^T): db must be ZERO on entry; xdot_scratch: C*L floats. */
int ssr_head_loss_bnsums(const float* feat, const float* feat_stats, const float* w, const float* bias, const float* image,
                         int image_channels, const int* res_idx, const float* target, float* pred, float* dfeat,
                         float* dw, float* db, double* loss, float* gout_scratch, int B, int d0, int d1, int d2, int C,
                         int L, int metric, const int* crop_size, const int* crop_begin, float* xdot_scratch,
                         double* sums2, void* stream);
/* keras.optimizers.Adam (SynthSR/training.py:444) on one flat parameter buffer. */
int ssr_adam_flat(float* p, const float* g, float* m, float* v, long long n, float lr_t, float beta1, float beta2,
                  float eps, float grad_scale, void* stream);

/* ---------------------------------------------------------------- segmentation-regularised loss (8f rank 4) -- */
/* SynthSR/metrics_model.py:136-215 (add_seg_loss_to_model) + ext/lab2im/layers.py:1334-1376 (DiceLoss, enable_checks=False).
 * NOT YET VALIDATED ON A B200; oracle: oracle/unet.py:seg_regularised_loss.
 * input of the frozen segmentation network: y = pred (+ image[..., res_channel]) [-> (clip(y, m, M) - m) / (M - m)]  (:151-154) */
int ssr_seg_input(const float* pred, const float* image, int image_channels, int res_channel, int use_clip, float m, float M,
                  float* y, long long nvox, void* stream);
int ssr_seg_input_bwd(const float* pred, const float* image, int image_channels, int res_channel, int use_clip, float m,
                      float M, const float* dy, float* dpred, long long nvox, void* stream);
/* softmax over S segmentation logits per voxel, channels merged into K generation classes (cls_of_seg[j] = class or -1,
 * device int[S]); gt of class k = [label == gt_value[k]] (device int[K]; the reference compares with the loop index, :188);
 * sums [B][K][2] doubles = sum 2 gt p | sum gt^2 + p^2 over the (optionally cropped, HOST int[3] size / begin) volume. */
int ssr_softmax_dice_sums(const float* logits, int S, const int* labels, const int* cls_of_seg, const int* gt_value, int K,
                          int B, int d0, int d1, int d2, const int* crop_size, const int* crop_begin, double* sums,
                          void* stream);
/* loss[0] += rel_weight * mean_{b,k} (1 - (top + 1e-7) / (bottom + 1e-7))   (layers.py:1363-1376, metrics_model.py:209) */
int ssr_dice_finalize(const double* sums, int B, int K, double rel_weight, double* loss, void* stream);
/* gradient of rel_weight * dice loss w.r.t. the logits (zero outside the crop) */
int ssr_softmax_dice_grad(const float* logits, int S, const int* labels, const int* cls_of_seg, const int* gt_value, int K,
                          int B, int d0, int d1, int d2, const int* crop_size, const int* crop_begin, const double* sums,
                          float rel_weight, float* dlogits, void* stream);
/* extra gradient e = dL_dice/d(prediction) [nvox] through the main network's single-output 1x1x1 head: dfeat += e w^T,
 * dw += feat^T e, db += sum e; feat_stats (optional): the folded BatchNorm of ssr_head_loss (x * s[2C+c] + s[3C+c]). */
int ssr_head_extra_grad(const float* feat, const float* feat_stats, const float* w, const float* e, long long nvox, int C,
                        float* dfeat, float* dw, float* db, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SYNTHSR_B200_H_ */
