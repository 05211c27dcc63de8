// sampler_api.cu -- C-ABI entry points of the spatial correlation sampler (include/b200corr.h).
//
// Replaces the pybind11 `forward` / `backward` of the reference's
// spatial_correlation_sampler_backend (correlation_sampler.cpp:59-129): same 12 integers in the
// same order, same output shape rule, but plain device pointers, an explicit stream, caller-owned
// outputs and an integer error code instead of a C++ exception.  Dispatch: the register-blocked
// TMA kernels of sampler_fast.cu when the problem has the FlowNetC / PWC-Net call-site structure
// (kernel 1, stride 1, padding 0, dilation 1, fp32), otherwise the generic kernels of
// sampler_generic.cu.  There is no CPU branch (the reference's correlation_sampler.cpp:76-86 has one).
#include "common.cuh"

namespace {

int check_common(const char *who, int B, int C, int H, int W, const int *q, int dtype, int *oH,
                 int *oW) {
  B200_CHECK(dtype >= B200CORR_F32 && dtype <= B200CORR_BF16, "%s: unsupported dtype %d", who, dtype);
  B200_CHECK(B >= 0 && C >= 0 && H >= 0 && W >= 0, "%s: negative tensor size", who);
  B200_CHECK(q[0] >= 1 && q[1] >= 1 && q[2] >= 1 && q[3] >= 1, "%s: kernel/patch size must be >= 1",
             who);
  B200_CHECK(q[4] >= 0 && q[5] >= 0, "%s: negative padding", who);
  B200_CHECK(q[6] >= 1 && q[7] >= 1 && q[8] >= 1 && q[9] >= 1 && q[10] >= 1 && q[11] >= 1,
             "%s: dilation / stride must be >= 1", who);
  *oH = b200corr_sampler_out_size(H, q[4], q[0], q[6], q[10]);
  *oW = b200corr_sampler_out_size(W, q[5], q[1], q[7], q[11]);
  B200_CHECK(*oH >= 0 && *oW >= 0, "%s: kernel does not fit the padded input", who);
  return 0;
}

}  // namespace

extern "C" {

int b200corr_sampler_out_size(int in_size, int pad, int kernel, int dilation, int stride) {
  const int span = in_size + 2 * pad - ((kernel - 1) * dilation + 1);
  if (span < 0) return -1;
  return span / stride + 1;
}

size_t b200corr_sampler_forward_workspace_bytes(int, int, int, int, int, int, int, int, int, int,
                                                int, int, int, int, int, int, int) {
  return 0;
}
size_t b200corr_sampler_backward_workspace_bytes(int B, int C, int H, int W, int kH, int kW,
                                                 int patchH, int patchW, int padH, int padW,
                                                 int dilationH, int dilationW, int dilation_patchH,
                                                 int dilation_patchW, int dH, int dW, int dtype) {
  const int q[12] = {kH, kW, patchH, patchW, padH, padW, dilationH, dilationW,
                     dilation_patchH, dilation_patchW, dH, dW};
  if (!b200::sampler_fast_applicable(B, C, H, W, q, dtype, 1)) return 0;
  return sizeof(int) * b200::sampler_fast_backward_plan_ints(B, C, H, W, q);
}

int b200corr_sampler_backward_plan(int B, int C, int H, int W, int kH, int kW, int patchH,
                                   int patchW, int padH, int padW, int dilationH, int dilationW,
                                   int dilation_patchH, int dilation_patchW, int dH, int dW,
                                   int dtype, void *h_plan, size_t bytes) {
  const int q[12] = {kH, kW, patchH, patchW, padH, padW, dilationH, dilationW,
                     dilation_patchH, dilation_patchW, dH, dW};
  if (!b200::sampler_fast_applicable(B, C, H, W, q, dtype, 1)) return 0;
  B200_CHECK(h_plan, "sampler_backward_plan: null buffer");
  return b200::sampler_fast_backward_plan(B, C, H, W, q, (int *)h_plan, bytes);
}

int b200corr_sampler_uses_fast_path(int B, int C, int H, int W, int kH, int kW, int patchH,
                                    int patchW, int padH, int padW, int dilationH, int dilationW,
                                    int dilation_patchH, int dilation_patchW, int dH, int dW,
                                    int dtype, int backward) {
  const int q[12] = {kH, kW, patchH, patchW, padH, padW, dilationH, dilationW,
                     dilation_patchH, dilation_patchW, dH, dW};
  return b200::sampler_fast_applicable(B, C, H, W, q, dtype, backward) ? 1 : 0;
}

int b200corr_sampler_forward(const void *in1, const void *in2, void *out, void *workspace,
                             size_t workspace_bytes, int B, int C, int H, int W, int kH, int kW,
                             int patchH, int patchW, int padH, int padW, int dilationH,
                             int dilationW, int dilation_patchH, int dilation_patchW, int dH, int dW,
                             int dtype, void *stream_) {
  (void)workspace;
  (void)workspace_bytes;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int q[12] = {kH, kW, patchH, patchW, padH, padW, dilationH, dilationW,
                     dilation_patchH, dilation_patchW, dH, dW};
  int oH, oW;
  if (int e = check_common("sampler_forward", B, C, H, W, q, dtype, &oH, &oW)) return e;
  if ((size_t)B * patchH * patchW * oH * oW == 0) return 0;
  B200_CHECK(in1 && in2 && out, "sampler_forward: null pointer");
  if (b200::sampler_fast_applicable(B, C, H, W, q, dtype, 0))
    return b200::sampler_fast_forward((const float *)in1, (const float *)in2, (float *)out, B, C, H,
                                      W, q, stream);
  return b200::sampler_generic_forward(in1, in2, out, B, C, H, W, oH, oW, q, dtype, stream);
}

int b200corr_sampler_backward(const void *in1, const void *in2, const void *grad_out,
                              void *grad_in1, void *grad_in2, void *workspace,
                              size_t workspace_bytes, int B, int C, int H, int W, int kH, int kW,
                              int patchH, int patchW, int padH, int padW, int dilationH,
                              int dilationW, int dilation_patchH, int dilation_patchW, int dH,
                              int dW, int dtype, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int q[12] = {kH, kW, patchH, patchW, padH, padW, dilationH, dilationW,
                     dilation_patchH, dilation_patchW, dH, dW};
  int oH, oW;
  if (int e = check_common("sampler_backward", B, C, H, W, q, dtype, &oH, &oW)) return e;
  if ((size_t)B * C * H * W == 0) return 0;
  B200_CHECK(in1 && in2 && grad_in1 && grad_in2, "sampler_backward: null pointer");
  B200_CHECK(grad_out || (size_t)patchH * patchW * oH * oW == 0, "sampler_backward: null grad_out");
  if (b200::sampler_fast_applicable(B, C, H, W, q, dtype, 1)) {
    // workspace (optional): the device copy of b200corr_sampler_backward_plan()'s schedule
    const size_t need = sizeof(int) * b200::sampler_fast_backward_plan_ints(B, C, H, W, q);
    const int *plan = (workspace && need > 0 && workspace_bytes >= need) ? (const int *)workspace : nullptr;
    return b200::sampler_fast_backward((const float *)in1, (const float *)in2,
                                       (const float *)grad_out, (float *)grad_in1,
                                       (float *)grad_in2, B, C, H, W, q, plan, stream);
  }
  return b200::sampler_generic_backward(in1, in2, grad_out, grad_in1, grad_in2, B, C, H, W, oH, oW,
                                        q, dtype, stream);
}

}  // extern "C"
