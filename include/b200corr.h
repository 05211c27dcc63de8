/*
 * b200corr.h -- C ABI of libb200corr.so: the B200 (sm_100a) cost-volume correlation hot path.
 *
 * This is the drop-in boundary.  Each entry point replaces one binding of the reference's two
 * pybind11 extensions (citations relative to /root/reference):
 *
 *   b200corr_sampler_forward   <- spatial_correlation_sampler_backend.forward
 *                                 models/Pytorch-Correlation-extension/Correlation_Module/
 *                                 correlation_sampler.cpp:59-87,127 (CUDA: correlation_cuda_kernel.cu:236-273)
 *   b200corr_sampler_backward  <- spatial_correlation_sampler_backend.backward
 *                                 correlation_sampler.cpp:89-124,128 (CUDA: correlation_cuda_kernel.cu:275-327)
 *   b200corr_altcorr_forward   <- alt_cuda_corr.forward   models/alt_cuda_corr/correlation.cpp:23-32,52
 *                                 (kernel: correlation_kernel.cu:18-119, 260-286)
 *   b200corr_altcorr_backward  <- alt_cuda_corr.backward  models/alt_cuda_corr/correlation.cpp:35-48,53
 *                                 (kernel: correlation_kernel.cu:122-256, 288-324)
 *   b200corr_allpairs_pyramid  <- CorrBlock.__init__ / CorrBlock.corr, models/raft/corr.py:55-64,98-106
 *                                 (torch.matmul + "/ sqrt(dim)" + 3x F.avg_pool2d)
 *   b200corr_lookup_forward    <- CorrBlock.__call__, models/raft/corr.py:72-96 +
 *                                 bilinear_sampler, models/raft/utils/utils.py:62-76 (F.grid_sample)
 *   b200corr_lookup_backward,
 *   b200corr_pyramid_backward  <- what autograd derives for corr.py:62-64,72-96
 *
 * Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer on the current CUDA device
 *     unless its name starts with `h_` (host array);
 *   - all tensors are dense row-major ("contiguous" in the reference's CHECK_CONTIGUOUS sense);
 *   - the library never allocates or frees device memory: outputs and scratch are provided by the
 *     caller, scratch sizes come from the *_workspace_bytes() queries (host-only, no GPU needed);
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*) and the call returns
 *     immediately; the reference launches on the legacy default stream instead;
 *   - return value 0 = success; <0 = error, message via b200corr_last_error() (thread-local).
 *     The reference raises a C++ exception through TORCH_CHECK; the Python host layer turns the
 *     error code back into RuntimeError;
 *   - there is no CPU implementation behind any of these symbols.
 */
#ifndef B200CORR_H_
#define B200CORR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200CORR_VERSION 100

/* element types of the sampler entry points (the reference dispatches float/double on CPU and
 * float/double/half on CUDA, correlation.cpp:104, correlation_cuda_kernel.cu:262) */
#define B200CORR_F32 0
#define B200CORR_F64 1
#define B200CORR_F16 2   /* generic kernels only: half storage, fp32 arithmetic */
#define B200CORR_BF16 3  /* generic kernels only: bfloat16 storage, fp32 arithmetic */

/* precision of the all-pairs contraction */
#define B200CORR_PREC_TF32 0   /* one tcgen05 kind::tf32 pass, inputs rounded to TF32 (rna)      */
#define B200CORR_PREC_TF32X3 1 /* split TF32: lo*hi + hi*lo + hi*hi on the tensor cores, fp32-level accuracy */
#define B200CORR_PREC_FP32 2   /* exact fp32 FMA accumulation on the CUDA cores (SIMT tile GEMM) */

/* coordinate arithmetic of the lookup */
#define B200CORR_LOOKUP_GRIDSAMPLE 0 /* reproduce utils.py:65-70 + grid_sample's un-normalise     */
#define B200CORR_LOOKUP_DIRECT 1     /* sample at the pixel coordinate directly (no round trip)   */

int b200corr_version(void);
const char *b200corr_last_error(void);

/* ---------------------------------------------------------------- spatial correlation sampler */

/* correlation.cpp:90-94 / correlation_cuda_kernel.cu:249-253 */
int b200corr_sampler_out_size(int in_size, int pad, int kernel, int dilation, int stride);

/* Scratch needed by forward / backward for this problem (0 is a valid answer). */
size_t b200corr_sampler_forward_workspace_bytes(int B, int C, int H, int W, int kH, int kW,
                                                int patchH, int patchW, int padH, int padW,
                                                int dilationH, int dilationW, int dilation_patchH,
                                                int dilation_patchW, int dH, int dW, int dtype);
size_t b200corr_sampler_backward_workspace_bytes(int B, int C, int H, int W, int kH, int kW,
                                                 int patchH, int patchW, int padH, int padW,
                                                 int dilationH, int dilationW, int dilation_patchH,
                                                 int dilation_patchW, int dH, int dW, int dtype);

/* Backward schedule (host only, no GPU needed): fills `h_plan` (HOST memory, at least
 * b200corr_sampler_backward_workspace_bytes() bytes) with a longest-processing-time-first
 * assignment of the backward work units to the persistent CTAs.  The caller copies it to the device
 * once per problem shape and passes that copy as `workspace` to b200corr_sampler_backward; without
 * it the kernels fall back to a static round-robin (same results, ~10% slower on ragged work).
 * The reference has no counterpart (it launches one kernel per batch sample,
 * correlation_cuda_kernel.cu:305-323). */
int b200corr_sampler_backward_plan(int B, int C, int H, int W, int kH, int kW, int patchH,
                                   int patchW, int padH, int padW, int dilationH, int dilationW,
                                   int dilation_patchH, int dilation_patchW, int dH, int dW,
                                   int dtype, void *h_plan, size_t bytes);

/* out[B, patchH, patchW, oH, oW] = sum_{c,i,j} in1[b,c,y1,x1] * in2[b,c,y1+dy,x1+dx]
 * (SURVEY.md section 8(0) S1).  Every element of `out` is written.  The 12 integers are in the
 * order of the reference's backend.forward(...) call (spatial_correlation_sampler.py:68-83). */
int b200corr_sampler_forward(const void *in1, const void *in2, void *out, void *workspace,
                             size_t workspace_bytes, int B, int C, int H, int W, int kH, int kW,
                             int patchH, int patchW, int padH, int padW, int dilationH,
                             int dilationW, int dilation_patchH, int dilation_patchW, int dH, int dW,
                             int dtype, void *stream);

/* grad_in1, grad_in2 [B, C, H, W] from grad_out [B, patchH, patchW, oH, oW]; every element of both
 * gradients is written (gather form, no atomics, deterministic). */
int b200corr_sampler_backward(const void *in1, const void *in2, const void *grad_out,
                              void *grad_in1, void *grad_in2, void *workspace,
                              size_t workspace_bytes, int B, int C, int H, int W, int kH, int kW,
                              int patchH, int patchW, int padH, int padW, int dilationH,
                              int dilationW, int dilation_patchH, int dilation_patchW, int dH,
                              int dW, int dtype, void *stream);

/* 1 if the problem runs on the TMA-staged register-blocked kernels, 0 if on the generic kernels. */
int b200corr_sampler_uses_fast_path(int B, int C, int H, int W, int kH, int kW, int patchH,
                                    int patchW, int padH, int padW, int dilationH, int dilationW,
                                    int dilation_patchH, int dilation_patchW, int dH, int dW,
                                    int dtype, int backward);

/* ---------------------------------------------------------------- fused FlowNetC merge block
 * (SURVEY.md section 8(f) row 2).  Replaces, at the FlowNetC call sites, the chain
 *   correlate(a, b) = spatial_correlation_sample(a, b, 1, patch, 1, 0, dilation_patch).view(B, P*P, H, W) / C
 *                                                                     models/submodules.py:124-138
 *   LeakyReLU(slope) ; torch.cat((conv_redir(a), out_corr), 1)       models/FlowNetC.py:133-147
 * by one kernel that writes leaky_relu(corr / C) into channels [c_off, c_off + P*P) of the caller's
 * (B, c_total, H, W) concat tensor `merged` (the other channels are not touched).  kernel 1, stride 1,
 * padding 0, fp32; only shapes b200corr_merge_supported() accepts (the register-blocked kernels). */
int b200corr_merge_supported(int B, int C, int H, int W, int patch, int dilation_patch);
int b200corr_merge_forward(const float *in1, const float *in2, float *merged, int B, int C, int H, int W,
                           int patch, int dilation_patch, int c_total, int c_off, float slope, void *stream);

/* Gradients w.r.t. in1 / in2 from the gradient of the concat tensor (same (B, c_total, H, W) layout) and
 * the forward's `merged` (its sign is the LeakyReLU mask).  grad_scratch: b200corr_merge_backward_scratch_bytes()
 * bytes of device memory (the masked, scaled, gathered slice); plan_workspace as in b200corr_sampler_backward. */
size_t b200corr_merge_backward_scratch_bytes(int B, int H, int W, int patch);
int b200corr_merge_backward(const float *in1, const float *in2, const float *merged, const float *grad_merged,
                            float *grad_in1, float *grad_in2, void *grad_scratch, void *plan_workspace,
                            size_t plan_workspace_bytes, int B, int C, int H, int W, int patch, int dilation_patch,
                            int c_total, int c_off, float slope, void *stream);

/* ---------------------------------------------------------------- PWC-Net warp
 * out[b,c,y,x] = mask * bilinear(in[b,c], position of (x + flow[b,0,y,x], y + flow[b,1,y,x]))  -- PWCDCNet.warp,
 * models/PWCNet.py:164-204: grid_sample (bilinear, zeros padding, align_corners=False on a grid normalised with
 * W-1 / H-1, exactly as the reference calls it) of the map and of an all-ones map, mask = [ones-sample >= 1e-4].
 * in / out (B,C,H,W), flow (B,2,H,W), fp32.  backward zero-fills and accumulates grad_in (atomics, as
 * grid_sample's own backward) and grad_flow; the mask passes no gradient. */
int b200corr_warp_forward(const float *in, const float *flow, float *out, int B, int C, int H, int W, void *stream);
int b200corr_warp_backward(const float *in, const float *flow, const float *grad_out, float *grad_in, float *grad_flow,
                           int B, int C, int H, int W, void *stream);

/* ---------------------------------------------------------------- FlowNet2 natives
 * Replace the reference's channelnorm_cuda / resample2d_cuda extensions
 * (models/channelnorm_package/channelnorm_cuda.cc:6-27, models/resample2d_package/resample2d_cuda.cc:6-25): fp32, dense
 * NCHW, outputs caller-owned.  norm_deg must be 2 and kernel_size 1 (what the reference uses; its kernels ignore
 * norm_deg and read out of bounds for larger kernel sizes).  resample2d clamps the sampling corners to the image. */
int b200corr_channelnorm_forward(const float *in, float *out, int B, int C, int H, int W, int norm_deg, void *stream);
int b200corr_channelnorm_backward(const float *in, const float *out, const float *grad_out, float *grad_in, int B, int C,
                                  int H, int W, int norm_deg, void *stream);
int b200corr_resample2d_forward(const float *in, const float *flow, float *out, int B, int C, int H, int W,
                                int kernel_size, int bilinear, void *stream);
int b200corr_resample2d_backward(const float *in, const float *flow, const float *grad_out, float *grad_in,
                                 float *grad_flow, int B, int C, int H, int W, int kernel_size, int bilinear, void *stream);

/* ---------------------------------------------------------------- RAFT CorrBlock */

/* Level l has shape (B*H*W, 1, H_l, W_l), H_0 = H, H_{l+1} = H_l / 2 (floor), same for W. */
size_t b200corr_allpairs_workspace_bytes(int B, int C, int H, int W, int precision);

/* h_levels[l] = device pointer of level l (l < num_levels <= 8).
 * level0[b,p1,p2] = scale * sum_c f1[b,c,p1] * f2[b,c,p2]; level l+1 = avg_pool2d(level l, 2, 2). */
int b200corr_allpairs_pyramid(const float *f1, const float *f2, float *const *h_levels,
                              int num_levels, int B, int C, int H, int W, float scale,
                              int precision, void *workspace, size_t workspace_bytes, void *stream);

/* Same with maps of different sizes: f1 [B,C,H1,W1] (queries), f2 [B,C,H2,W2] (keys); level l has shape
 * (B*H1*W1, 1, H2_l, W2_l).  What AlternateCorrBlock needs for its pooled key maps (corr.py:114-118).
 * Tensor-core precisions only (TF32, TF32X3), W2 % 4 == 0. */
size_t b200corr_allpairs_rect_workspace_bytes(int B, int C, int H1, int W1, int H2, int W2, int precision);
int b200corr_allpairs_pyramid_rect(const float *f1, const float *f2, float *const *h_levels,
                                   int num_levels, int B, int C, int H1, int W1, int H2, int W2, float scale,
                                   int precision, void *workspace, size_t workspace_bytes, void *stream);

/* Blocked volume layout (B200-first memory layout; no counterpart in the reference).  A blocked level stores each
 * query's slice as a row-major grid of 8x8 tiles of 64 consecutive floats, padded to whole tiles:
 *     offset(y, x) = ((y / 8) * (Wp / 8) + x / 8) * 64 + (y % 8) * 8 + x % 8,   slice = Hp * Wp floats
 * with (Hp, Wp) from b200corr_blocked_level_dims (level 0: H rounded up to 8, W; level 1: rounded up to whole tiles).
 * Padding inside written tiles holds zeros (= the zero padding of grid_sample); rows past the level's extent are
 * never read.  The 10-row windows of the lookup then touch 4-6 tiles of 256 B instead of 10 rows W_l * 4 bytes apart,
 * and the volume is written in aligned 1 KB runs whatever W is (row-major rows of 480 B, e.g. 68x120 maps, make every
 * other 128-byte store straddle two lines shared with another CTA).  Bit l of `blocked_levels` = level l is blocked;
 * b200corr_allpairs_blocked_levels() returns the mask this build supports for a problem (levels 0 and 1 when
 * W % 8 == 0 and the tensor-core kernels run as CTA pairs, else 0); pass that mask or 0.  The same mask goes to
 * b200corr_lookup_forward_layout.  The reference's observable (B*H*W, 1, H_l, W_l) row-major pyramid (corr.py:66-67)
 * is what mask 0 produces. */
void b200corr_blocked_level_dims(int level, int H, int W, int *Hp, int *Wp);
int b200corr_allpairs_blocked_levels(int num_levels, int H, int W, int precision);
int b200corr_allpairs_pyramid_layout(const float *f1, const float *f2, float *const *h_levels,
                                     int num_levels, int B, int C, int H1, int W1, int H2, int W2, float scale,
                                     int precision, int blocked_levels, void *workspace, size_t workspace_bytes,
                                     void *stream);
int b200corr_lookup_forward_layout(const float *const *h_levels, int num_levels, int first_level, int blocked_levels,
                                   const float *coords, float *out, int B, int H, int W, int radius, int mode,
                                   void *stream);

/* fp16 storage of the blocked levels (opt-in; no counterpart in the reference, whose volume is fp32).  Bit l of
 * `half_levels` = level l holds fp16 values in the same tile order (a tile is 64 halves = 128 B, offsets and slice
 * sizes above count elements); half_levels must be 0 or equal to blocked_levels, h_levels[l] of such a level points
 * to the fp16 buffer.  Each value is rounded once from the fp32 accumulator / the fp32 2x2 mean (round to nearest,
 * relative 2^-11 -- the size of the TF32 input rounding; finite values beyond +-65504 saturate); the coarse levels
 * stay fp32.  Halves what the build writes and what the lookups read of the two levels that hold 94 % of the volume. */
int b200corr_allpairs_pyramid_storage(const float *f1, const float *f2, float *const *h_levels,
                                      int num_levels, int B, int C, int H1, int W1, int H2, int W2, float scale,
                                      int precision, int blocked_levels, int half_levels, void *workspace,
                                      size_t workspace_bytes, void *stream);
int b200corr_lookup_forward_storage(const float *const *h_levels, int num_levels, int first_level, int blocked_levels,
                                    int half_levels, const float *coords, float *out, int B, int H, int W, int radius,
                                    int mode, void *stream);

/* out[B, num_levels*(2r+1)^2, H, W]; coords[B, 2, H, W] (channel 0 = x).  Channel index
 * l*(2r+1)^2 + i*(2r+1) + j, i = x-offset index, j = y-offset index (corr.py:80-86). */
int b200corr_lookup_forward(const float *const *h_levels, int num_levels, const float *coords,
                            float *out, int B, int H, int W, int radius, int mode, void *stream);

/* Same for a run of pyramid levels that does not start at level 0: list entry i is pyramid level
 * first_level + i, i.e. has extent (H, W) >> (first_level + i) and samples at coords / 2^(first_level + i);
 * out[B, num_levels*(2r+1)^2, H, W] holds just these levels. */
int b200corr_lookup_forward_from(const float *const *h_levels, int num_levels, int first_level,
                                 const float *coords, float *out, int B, int H, int W, int radius, int mode,
                                 void *stream);

/* ---- lookup fused with the motion encoder's first convolution.
 * Reference: corr = corr_fn(coords1) (models/raft/raft.py:189) followed by cor = F.relu(self.convc1(corr))
 * (models/raft/update.py:104,111; convc1 = Conv2d(L*(2r+1)^2, n_out, 1)).  The (B, L*(2r+1)^2, H, W) lookup
 * result is never materialised: out[B, n_out, H, W] = act(sum_k W[:, k] * lookup[:, k] + bias), act = ReLU if
 * `relu`.  Samples and weights are rounded to TF32, accumulation in fp32 (torch's default for this convolution:
 * torch.backends.cudnn.allow_tf32).
 * prepare: weight [n_out, L*(2r+1)^2] (the conv weight with its 1x1 spatial dims dropped) -> wprep, once per
 * weight update; wprep holds b200corr_lookup_convc1_weight_bytes bytes, 16-byte aligned.
 * forward: levels / blocked_levels / coords / mode as b200corr_lookup_forward_layout (first_level 0);
 * n_out a multiple of 32, <= 256; bias [n_out] or NULL. */
size_t b200corr_lookup_convc1_weight_bytes(int num_levels, int radius, int n_out);
int b200corr_lookup_convc1_prepare(const float *weight, float *wprep, int num_levels, int radius, int n_out,
                                   void *stream);
int b200corr_lookup_convc1_forward(const float *const *h_levels, int num_levels, int blocked_levels,
                                   const float *coords, const float *wprep, const float *bias, float *out, int B,
                                   int H, int W, int radius, int mode, int n_out, int relu, void *stream);

/* The same convolution as a stand-alone tensor-core kernel behind the plain lookup ("pipelined" variant of the row
 * above; also usable for any 1x1 convolution): out[B, N, HW] = act(weight[N, K] . x[B, K, HW] + bias), TF32 tcgen05,
 * x read as it lies (MN-major operand through TMA).  N <= 256, K % 4 == 0, HW % 4 == 0, 16-byte aligned operands;
 * bias may be NULL; relu != 0 applies max(., 0). */
int b200corr_conv1x1_forward(const float *x, const float *weight, const float *bias, float *out, int B, int K,
                             int N, int HW, int relu, void *stream);

/* Accumulates (+=) d(out)/d(level l) into h_grad_levels[l] (caller zero-initialises once per
 * CorrBlock; several lookups of the same block accumulate).  Coordinates get no gradient
 * (the reference detaches them, models/raft/raft.py:188). */
int b200corr_lookup_backward(float *const *h_grad_levels, int num_levels, const float *coords,
                             const float *grad_out, int B, int H, int W, int radius, int mode,
                             void *stream);

/* Folds the pooled-level gradients down into level 0: for l = num_levels-1 .. 1,
 * grad_levels[l-1][.., y, x] += grad_levels[l][.., y/2, x/2] / 4 (backward of avg_pool2d). */
int b200corr_pyramid_backward(float *const *h_grad_levels, int num_levels, int B, int H, int W,
                              void *stream);

/* Backward of the volume + pyramid w.r.t. the feature maps (what autograd derives for corr.py:55-64,98-106):
 * h_grad_levels[l] = dL/d vol_l, (B*H*W, 1, H>>l, W>>l) row-major, as b200corr_lookup_backward leaves them
 * (NOT folded: do not call b200corr_pyramid_backward first).  Writes (overwrites)
 *   grad_fmap1 = scale * sum_l G_l . pool_l(fmap2)^T      grad_fmap2 = scale * sum_l unpool_l(G_l^T . fmap1) / 4^l
 * which equals scale * fold(G) . F2^T and scale * fold(G)^T . F1 because average pooling commutes with the
 * contraction.  precision B200CORR_PREC_FP32: exact fp32 (CUDA cores); otherwise TF32 tensor cores (operands
 * truncated to TF32, fp32 accumulation), falling back to the exact kernel for shapes TMA cannot express.
 * workspace: b200corr_volume_backward_workspace_bytes (pooled copies of fmap2 and their gradients). */
size_t b200corr_volume_backward_workspace_bytes(int num_levels, int B, int C, int H, int W);
int b200corr_volume_backward(const float *const *h_grad_levels, int num_levels, const float *fmap1,
                             const float *fmap2, float *grad_fmap1, float *grad_fmap2, int B, int C, int H, int W,
                             float scale, int precision, void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------- alt_cuda_corr */

/* fmap1 [B,H1,W1,C], fmap2 [B,H2,W2,C] (NHWC), coords [B,N,H1,W1,2] -> corr [B,N,(2r+1)^2,H1,W1],
 * channel = ix*(2r+1) + iy, no 1/sqrt(C) factor (correlation_kernel.cu:92-114; corr.py:137). */
int b200corr_altcorr_forward(const float *fmap1, const float *fmap2, const float *coords,
                             float *corr, int B, int N, int H1, int W1, int H2, int W2, int C,
                             int radius, void *stream);

/* fmap1_grad [B,H1,W1,C] and fmap2_grad [B,H2,W2,C] are fully written (fmap2_grad is zeroed
 * inside the call, then scatter-added); coords_grad [B,N,H1,W1,2] is set to zero, as in the
 * reference (correlation_kernel.cu:307). */
int b200corr_altcorr_backward(const float *fmap1, const float *fmap2, const float *coords,
                              const float *corr_grad, float *fmap1_grad, float *fmap2_grad,
                              float *coords_grad, int B, int N, int H1, int W1, int H2, int W2,
                              int C, int radius, void *stream);

/* ---------------------------------------------------------------- universal-patch placement */

/* The host-side transform of the patch attack on the GPU.  Reference: patch_attacks/utils_patch.py:257-358
 * (`circle_transform`: brightness jitter + clip, patch * mask, zoom, rotate, paste at a random location) and
 * patch_attacks/main.py:537-542 (`adv = (1 - mask_canvas) * img + patch_canvas`).
 * patch [3,p,p], mask [p,p], placements [n,5] = (scale, angle [rad], centre x, centre y, brightness offset),
 * img1 / img2 / adv1 / adv2 (and the gradients below) [n,3,H,W] sharing the ELEMENT strides sN, sC, sH, sW
 * (contiguous or channels_last):
 *   q = clamp(patch + bright, 0, 1) * mask;  (u, v) = R(-angle) (x - cx, y - cy) / scale + (p-1)/2
 *   adv_k = clamp((1 - bilinear(mask; u, v)) * img_k + bilinear(q; u, v), 0, 1)      zeros outside the patch */
int b200corr_patch_compose_forward(const float *img1, const float *img2, const float *patch, const float *mask,
                                   const float *placements, float *adv1, float *adv2, int n, int H, int W, int p,
                                   long long sN, long long sC, long long sH, long long sW, void *stream);

/* grad_patch [3,p,p] = d(<grad_adv1, adv1> + <grad_adv2, adv2>) / d patch, summed over the n pairs in index
 * order (gather form, deterministic, no atomics).  scratch: b200corr_patch_compose_backward_scratch_bytes. */
size_t b200corr_patch_compose_backward_scratch_bytes(int n, int p);
int b200corr_patch_compose_backward(const float *img1, const float *img2, const float *patch, const float *mask,
                                    const float *placements, const float *grad_adv1, const float *grad_adv2,
                                    float *grad_patch, float *scratch, size_t scratch_bytes, int n, int H, int W, int p,
                                    long long sN, long long sC, long long sH, long long sW, void *stream);

/* ---------------------------------------------------------------- diagnostics */

/* Runs `iters` dependent FP32 FMAs per thread on every SM and returns the achieved TFLOP/s in
 * *tflops (used by bench.py for the FP32-pipe roofline denominator). */
int b200corr_measure_fp32_peak(int iters, float *tflops, void *stream);

/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
uint64_t b200corr_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200CORR_H_ */
