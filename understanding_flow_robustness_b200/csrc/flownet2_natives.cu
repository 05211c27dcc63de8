// flownet2_natives.cu -- the two small native ops of the FlowNet2 family:
//
//   ChannelNorm  models/channelnorm_package/channelnorm_kernel.cu:19-96
//       out[b,0,y,x] = sqrt(sum_c in[b,c,y,x]^2);   gin[b,c,y,x] = gout[b,0,y,x] * in[b,c,y,x] / (out[b,0,y,x] + 1e-9)
//   Resample2d   models/resample2d_package/resample2d_kernel.cu:17-195 (kernel_size 1, the only value the reference uses)
//       out[b,c,y,x] = bilinear(in[b,c], x + flow[b,0,y,x], y + flow[b,1,y,x]) with the four corners CLAMPED to the
//       image (border replication, not grid_sample's zero padding); nearest neighbour when bilinear == 0.
//
// Glue around the correlation path (they act on 3-channel images and 2-channel flows), restated with the reference's
// own arithmetic so that FlowNet2 runs without the legacy -gencode builds of the reference extensions: the forward
// weights are formed in double and rounded per term (`(1. - alpha) * (1. - beta) * v`, :53-56), the backward uses
// `xf - int(xf)` (truncation, :96-97) where the forward uses floor -- both kept as they are.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
channelnorm_fwd_kernel(const float *__restrict__ in, float *__restrict__ out, int B, int C, long long HW) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * HW) return;
  const long long b = i / HW, r = i - b * HW;
  const float *p = in + b * C * HW + r;
  float acc = 0.f;
  for (int c = 0; c < C; ++c) {
    const float v = p[(long long)c * HW];
    acc = fmaf(v, v, acc);            // `result += val * val` under nvcc's default contraction
  }
  out[i] = sqrtf(acc);
}

__global__ void __launch_bounds__(256)
channelnorm_bwd_kernel(const float *__restrict__ in, const float *__restrict__ out, const float *__restrict__ gout,
                       float *__restrict__ gin, int B, int C, long long HW) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * C * HW) return;
  const long long b = i / (C * HW), r = i % HW;
  const long long o = b * HW + r;
  // float * float / (float + double 1e-9): the division runs in double, as written in the reference (:93)
  gin[i] = (float)((double)(gout[o] * in[i]) / ((double)out[o] + 1e-9));
}

struct Corners {
  int xL, xR, yT, yB;
};
__device__ __forceinline__ Corners corners(float xf, float yf, int H, int W) {
  Corners c;
  c.xL = max(min((int)floorf(xf), W - 1), 0);
  c.xR = max(min((int)(floorf(xf) + 1.f), W - 1), 0);
  c.yT = max(min((int)floorf(yf), H - 1), 0);
  c.yB = max(min((int)(floorf(yf) + 1.f), H - 1), 0);
  return c;
}

__global__ void __launch_bounds__(256)
resample2d_fwd_kernel(const float *__restrict__ in, const float *__restrict__ flow, float *__restrict__ out, int B, int C,
                      int H, int W, int bilinear) {
  const long long HW = (long long)H * W;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * C * HW) return;
  const int x = (int)(i % W), y = (int)((i / W) % H);
  const long long bc = i / HW, b = bc / C;
  const long long r = (long long)y * W + x;
  const float xf = (float)x + flow[(b * 2) * HW + r], yf = (float)y + flow[(b * 2 + 1) * HW + r];
  const float *p = in + bc * HW;
  if (bilinear) {
    const float alpha = xf - floorf(xf), beta = yf - floorf(yf);
    const Corners c = corners(xf, yf, H, W);
    float val = 0.f;
    val += (float)((1. - alpha) * (1. - beta) * p[(long long)c.yT * W + c.xL]);
    val += (float)((alpha) * (1. - beta) * p[(long long)c.yT * W + c.xR]);
    val += (float)((1. - alpha) * (beta) * p[(long long)c.yB * W + c.xL]);
    val += (float)((alpha) * (beta) * p[(long long)c.yB * W + c.xR]);
    out[i] = val;
  } else {
    // floor(x + 0.5) in double, as resample2d_kernel.cu:52-53 computes it (an fp32 add can round up to the next integer)
    const int xN = max(min((int)floor((double)xf + 0.5), W - 1), 0), yN = max(min((int)floor((double)yf + 0.5), H - 1), 0);
    out[i] = p[(long long)yN * W + xN];
  }
}

// grad w.r.t. the image: scatter with atomics (resample2d_kernel.cu:64-124); gin zero-initialised by the caller
__global__ void __launch_bounds__(256)
resample2d_bwd_in_kernel(const float *__restrict__ flow, const float *__restrict__ gout, float *__restrict__ gin, int B,
                         int C, int H, int W) {
  const long long HW = (long long)H * W;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * C * HW) return;
  const int x = (int)(i % W), y = (int)((i / W) % H);
  const long long bc = i / HW, b = bc / C;
  const long long r = (long long)y * W + x;
  const float xf = (float)x + flow[(b * 2) * HW + r], yf = (float)y + flow[(b * 2 + 1) * HW + r];
  const float alpha = xf - (float)(int)xf, beta = yf - (float)(int)yf;   // truncation, as the reference
  const Corners c = corners(xf, yf, H, W);
  const float g = gout[i];
  float *q = gin + bc * HW;
  atomicAdd(q + (long long)c.yT * W + c.xL, (1 - alpha) * (1 - beta) * g);
  atomicAdd(q + (long long)c.yT * W + c.xR, (alpha) * (1 - beta) * g);
  atomicAdd(q + (long long)c.yB * W + c.xL, (1 - alpha) * (beta) * g);
  atomicAdd(q + (long long)c.yB * W + c.xR, (alpha) * (beta) * g);
}

// grad w.r.t. the flow (resample2d_kernel.cu:126-195): channel 0 differentiates along x, channel 1 along y
__global__ void __launch_bounds__(256)
resample2d_bwd_flow_kernel(const float *__restrict__ in, const float *__restrict__ flow, const float *__restrict__ gout,
                           float *__restrict__ gflow, int B, int C, int H, int W) {
  const long long HW = (long long)H * W;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * 2 * HW) return;
  const int x = (int)(i % W), y = (int)((i / W) % H);
  const int ch2 = (int)((i / HW) % 2);
  const long long b = i / (2 * HW);
  const long long r = (long long)y * W + x;
  const float xf = (float)x + flow[(b * 2) * HW + r], yf = (float)y + flow[(b * 2 + 1) * HW + r];
  const Corners c = corners(xf, yf, H, W);
  float o = 0.f;
  if (ch2 % 2) {
    const float gamma = 1 - (xf - floorf(xf));
    for (int ch = 0; ch < C; ++ch) {
      const float *p = in + (b * C + ch) * HW;
      const float g = gout[(b * C + ch) * HW + r];
      o += (gamma)*g * p[(long long)c.yB * W + c.xL];
      o -= (gamma)*g * p[(long long)c.yT * W + c.xL];
      o += (1 - gamma) * g * p[(long long)c.yB * W + c.xR];
      o -= (1 - gamma) * g * p[(long long)c.yT * W + c.xR];
    }
  } else {
    const float gamma = 1 - (yf - floorf(yf));
    for (int ch = 0; ch < C; ++ch) {
      const float *p = in + (b * C + ch) * HW;
      const float g = gout[(b * C + ch) * HW + r];
      o += (gamma)*g * p[(long long)c.yT * W + c.xR];
      o -= (gamma)*g * p[(long long)c.yT * W + c.xL];
      o += (1 - gamma) * g * p[(long long)c.yB * W + c.xR];
      o -= (1 - gamma) * g * p[(long long)c.yB * W + c.xL];
    }
  }
  gflow[i] = o;
}

int check4(const char *who, int B, int C, int H, int W) {
  B200_CHECK(B >= 0 && C >= 1 && H >= 1 && W >= 1, "%s: bad sizes", who);
  return 0;
}
inline unsigned nblocks(long long n) { return (unsigned)((n + 255) / 256); }

}  // namespace

extern "C" {

int b200corr_channelnorm_forward(const float *in, float *out, int B, int C, int H, int W, int norm_deg, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int e = check4("channelnorm_forward", B, C, H, W)) return e;
  B200_CHECK(norm_deg == 2, "channelnorm_forward: norm_deg %d (the reference kernel computes the 2-norm whatever it is given)",
             norm_deg);
  if (B == 0) return 0;
  B200_CHECK(in && out, "channelnorm_forward: null pointer");
  const long long HW = (long long)H * W;
  channelnorm_fwd_kernel<<<nblocks(B * HW), 256, 0, stream>>>(in, out, B, C, HW);
  B200_LAUNCH_OK("channelnorm_fwd_kernel");
  return 0;
}

int b200corr_channelnorm_backward(const float *in, const float *out, const float *grad_out, float *grad_in, int B, int C,
                                  int H, int W, int norm_deg, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int e = check4("channelnorm_backward", B, C, H, W)) return e;
  B200_CHECK(norm_deg == 2, "channelnorm_backward: norm_deg %d", norm_deg);
  if (B == 0) return 0;
  B200_CHECK(in && out && grad_out && grad_in, "channelnorm_backward: null pointer");
  const long long HW = (long long)H * W;
  channelnorm_bwd_kernel<<<nblocks(B * C * HW), 256, 0, stream>>>(in, out, grad_out, grad_in, B, C, HW);
  B200_LAUNCH_OK("channelnorm_bwd_kernel");
  return 0;
}

int b200corr_resample2d_forward(const float *in, const float *flow, float *out, int B, int C, int H, int W,
                                int kernel_size, int bilinear, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int e = check4("resample2d_forward", B, C, H, W)) return e;
  B200_CHECK(kernel_size == 1, "resample2d_forward: kernel_size %d (the reference only ever uses 1; larger kernels read "
                               "past the clamped corners there)", kernel_size);
  if (B == 0) return 0;
  B200_CHECK(in && flow && out, "resample2d_forward: null pointer");
  resample2d_fwd_kernel<<<nblocks((long long)B * C * H * W), 256, 0, stream>>>(in, flow, out, B, C, H, W, bilinear);
  B200_LAUNCH_OK("resample2d_fwd_kernel");
  return 0;
}

int b200corr_resample2d_backward(const float *in, const float *flow, const float *grad_out, float *grad_in,
                                 float *grad_flow, int B, int C, int H, int W, int kernel_size, int bilinear,
                                 void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  (void)bilinear;   // the reference's backward uses the bilinear weights in either mode
  if (int e = check4("resample2d_backward", B, C, H, W)) return e;
  B200_CHECK(kernel_size == 1, "resample2d_backward: kernel_size %d", kernel_size);
  if (B == 0) return 0;
  B200_CHECK(in && flow && grad_out && grad_in && grad_flow, "resample2d_backward: null pointer");
  const long long n = (long long)B * C * H * W;
  B200_CUDA(cudaMemsetAsync(grad_in, 0, sizeof(float) * n, stream));
  resample2d_bwd_in_kernel<<<nblocks(n), 256, 0, stream>>>(flow, grad_out, grad_in, B, C, H, W);
  B200_LAUNCH_OK("resample2d_bwd_in_kernel");
  resample2d_bwd_flow_kernel<<<nblocks((long long)B * 2 * H * W), 256, 0, stream>>>(in, flow, grad_out, grad_flow, B, C, H, W);
  B200_LAUNCH_OK("resample2d_bwd_flow_kernel");
  return 0;
}

}  // extern "C"
