// merge_block.cu -- C-ABI entry points of the fused FlowNetC merge block.
//
// The reference runs, per forward pass (models/FlowNetC.py:133-147, models/submodules.py:124-138):
//     out_corr = spatial_correlation_sample(a, b, 1, 21, 1, 0, dilation_patch=2)   # (B,21,21,H,W) written
//     out_corr = out_corr.view(B, 441, H, W) / C                                   # read + write
//     out_corr = LeakyReLU(0.1)(out_corr)                                          # read + write
//     in_conv3_1 = cat((conv_redir(a), out_corr), 1)                               # read + write
// i.e. seven passes over the 13.5 MB/sample cost volume, and seven more in the backward pass.  Here
// the forward kernel applies the scale and the activation to its register accumulators and stores
// them straight into the channel slice of the concat tensor (one pass); the backward folds LeakyReLU'
// and 1/C into one elementwise pass that also gathers the slice, then runs the sampler backward.
#include "common.cuh"

namespace {

// g[b, k, p] = (merged[b, c_off + k, p] > 0 ? gm : gm * slope) * inv_c,  gm = grad_merged[b, c_off + k, p]
// (what autograd derives for cat -> LeakyReLU -> "/ C"; torch's CUDA division by a Python scalar is a
// multiplication by the fp32 reciprocal, so is this).  VEC = 4: planes are 16-byte aligned.
template <int VEC>
__global__ void __launch_bounds__(256)
merge_grad_kernel(const float *__restrict__ merged, const float *__restrict__ grad_merged,
                  float *__restrict__ g, long long plane, long long bstride, long long total, float slope,
                  float inv_c) {
  // plane = P*P*H*W floats of one sample's slice (contiguous inside the concat tensor)
  const long long stride = (long long)gridDim.x * blockDim.x * VEC;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * VEC; i < total; i += stride) {
    const long long b = i / plane, r = i - b * plane;
    const long long src = b * bstride + r;
    if constexpr (VEC == 4) {
      const float4 m = *reinterpret_cast<const float4 *>(merged + src);
      const float4 gm = __ldcs(reinterpret_cast<const float4 *>(grad_merged + src));
      float4 o;
      o.x = (m.x > 0.f ? gm.x : gm.x * slope) * inv_c;
      o.y = (m.y > 0.f ? gm.y : gm.y * slope) * inv_c;
      o.z = (m.z > 0.f ? gm.z : gm.z * slope) * inv_c;
      o.w = (m.w > 0.f ? gm.w : gm.w * slope) * inv_c;
      *reinterpret_cast<float4 *>(g + i) = o;
    } else {
      const float m = merged[src], gm = grad_merged[src];
      g[i] = (m > 0.f ? gm : gm * slope) * inv_c;
    }
  }
}

int merge_hyper(const char *who, int B, int C, int H, int W, int patch, int dilation_patch, int c_total,
                int c_off, float slope, int q[12]) {
  B200_CHECK(B >= 0 && C >= 1 && H >= 1 && W >= 1, "%s: bad sizes", who);
  B200_CHECK(patch >= 1 && dilation_patch >= 1, "%s: bad patch / dilation_patch", who);
  B200_CHECK(c_off >= 0 && c_total >= c_off + patch * patch, "%s: the %d correlation channels do not fit at offset %d of %d",
             who, patch * patch, c_off, c_total);
  B200_CHECK(slope >= 0.f, "%s: negative_slope must be >= 0 (the mask is recovered from the sign of the output)", who);
  const int qq[12] = {1, 1, patch, patch, 0, 0, 1, 1, dilation_patch, dilation_patch, 1, 1};
  for (int i = 0; i < 12; ++i) q[i] = qq[i];
  return 0;
}

}  // namespace

extern "C" {

int b200corr_merge_supported(int B, int C, int H, int W, int patch, int dilation_patch) {
  const int q[12] = {1, 1, patch, patch, 0, 0, 1, 1, dilation_patch, dilation_patch, 1, 1};
  return b200::sampler_fast_applicable(B, C, H, W, q, B200CORR_F32, 0) ? 1 : 0;
}

int b200corr_merge_forward(const float *in1, const float *in2, float *merged, int B, int C, int H, int W,
                           int patch, int dilation_patch, int c_total, int c_off, float slope, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int q[12];
  if (int e = merge_hyper("merge_forward", B, C, H, W, patch, dilation_patch, c_total, c_off, slope, q)) return e;
  if (B == 0) return 0;
  B200_CHECK(in1 && in2 && merged, "merge_forward: null pointer");
  B200_CHECK(b200::sampler_fast_applicable(B, C, H, W, q, B200CORR_F32, 0),
             "merge_forward: only the register-blocked structure is fused (patch 21 / dilation_patch 2 or patch 9 / 1, "
             "W %% 4 == 0); got patch %d dilation_patch %d C %d W %d", patch, dilation_patch, C, W);
  const long long HW = (long long)H * W;
  B200_CHECK((((uintptr_t)merged) & 15) == 0 && (HW * c_off) % 4 == 0 && (HW * c_total) % 4 == 0,
             "merge_forward: the slice must be 16-byte aligned");
  return b200::sampler_fast_forward_merge(in1, in2, merged + HW * c_off, B, C, H, W, q, HW * c_total, slope, stream);
}

size_t b200corr_merge_backward_scratch_bytes(int B, int H, int W, int patch) {
  return b200::align_up((size_t)B * patch * patch * H * W * sizeof(float), 256);
}

int b200corr_merge_backward(const float *in1, const float *in2, const float *merged, const float *grad_merged,
                            float *grad_in1, float *grad_in2, void *grad_scratch, void *plan_workspace,
                            size_t plan_workspace_bytes, int B, int C, int H, int W, int patch, int dilation_patch,
                            int c_total, int c_off, float slope, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int q[12];
  if (int e = merge_hyper("merge_backward", B, C, H, W, patch, dilation_patch, c_total, c_off, slope, q)) return e;
  if (B == 0) return 0;
  B200_CHECK(in1 && in2 && merged && grad_merged && grad_in1 && grad_in2 && grad_scratch, "merge_backward: null pointer");
  const long long HW = (long long)H * W, plane = HW * patch * patch, total = plane * B;
  const long long bstride = HW * c_total, off = HW * c_off;
  const bool vec = plane % 4 == 0 && bstride % 4 == 0 && off % 4 == 0 &&
                   ((((uintptr_t)merged) | ((uintptr_t)grad_merged) | ((uintptr_t)grad_scratch)) & 15) == 0;
  const float inv_c = 1.0f / (float)C;
  const long long work = vec ? total / 4 : total;
  long long blocks = (work + 255) / 256;
  const long long cap = (long long)b200::num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (vec)
    merge_grad_kernel<4><<<(unsigned)blocks, 256, 0, stream>>>(merged + off, grad_merged + off, (float *)grad_scratch, plane,
                                                               bstride, total, slope, inv_c);
  else
    merge_grad_kernel<1><<<(unsigned)blocks, 256, 0, stream>>>(merged + off, grad_merged + off, (float *)grad_scratch, plane,
                                                               bstride, total, slope, inv_c);
  B200_LAUNCH_OK("merge_grad_kernel");
  return b200corr_sampler_backward(in1, in2, grad_scratch, grad_in1, grad_in2, plan_workspace, plan_workspace_bytes, B, C,
                                   H, W, 1, 1, patch, patch, 0, 0, 1, 1, dilation_patch, dilation_patch, 1, 1,
                                   B200CORR_F32, stream_);
}

}  // extern "C"
