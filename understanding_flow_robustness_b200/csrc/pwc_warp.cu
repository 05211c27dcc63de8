// pwc_warp.cu -- PWC-Net's warp() in front of every correlation below the top level.
//
// Reference: models/PWCNet.py:164-204.  Per call it builds a mesh grid on the host, normalises
// grid + flow to [-1, 1], runs F.grid_sample on the feature map AND on an all-ones tensor of the same
// size (the validity mask), thresholds the mask at 1e-4 and multiplies: two gathers plus four
// elementwise passes over a (B, C, H, W) map.  Here: one gather kernel.  The sampling position, the four
// bilinear weights and the mask depend only on the pixel, so a thread computes them once and then walks
// its channels; the mask is the sum of the in-bounds weights, never materialised.
//
// Arithmetic follows the reference op by op in fp32 (grid_sample defaults: bilinear, zeros padding,
// align_corners=False -- the reference normalises for align_corners=True but calls the default, and so
// does this kernel):
//     v  = x + u                                  (PWCNet.py:184)
//     g  = (2 v) * (1 / max(W-1, 1)) - 1          (:189-190; torch divides by a Python scalar via the reciprocal)
//     ix = ((g + 1) * W - 1) / 2                  (ATen grid_sampler_unnormalize, align_corners=False)
//     out = sum_{corner in bounds} in[corner] * w_corner ;  mask = [sum_{corner in bounds} w_corner >= 1e-4]
#include "common.cuh"

namespace {

struct Taps {
  int x0, y0;           // north-west corner
  float w[4];           // nw, ne, sw, se weights, zero where the corner is outside the image
  float dwx[4], dwy[4]; // d(weight)/d(ix), d(weight)/d(iy) of the in-bounds corners
  float mask;           // 1 if the in-bounds weights sum to >= 1e-4
};

__device__ __forceinline__ Taps make_taps(float u, float v, int x, int y, int H, int W, float inv_w1, float inv_h1) {
  Taps t;
  const float vx = (float)x + u, vy = (float)y + v;
  const float gx = __fsub_rn(__fmul_rn(2.0f * vx, inv_w1), 1.0f);
  const float gy = __fsub_rn(__fmul_rn(2.0f * vy, inv_h1), 1.0f);
  const float ix = __fmaf_rn(gx + 1.0f, (float)W, -1.0f) * 0.5f;
  const float iy = __fmaf_rn(gy + 1.0f, (float)H, -1.0f) * 0.5f;
  const float fx = floorf(ix), fy = floorf(iy);
  // corners outside +-2^30 cannot be in bounds; clamp before the int conversion (NaN / inf flows)
  t.x0 = (int)fminf(fmaxf(fx, -1073741824.f), 1073741824.f);
  t.y0 = (int)fminf(fmaxf(fy, -1073741824.f), 1073741824.f);
  const float ax = ix - fx, ay = iy - fy;        // (ix - ix_nw), (iy - iy_nw)
  const float bx = (fx + 1.0f) - ix, by = (fy + 1.0f) - iy;   // (ix_se - ix), (iy_se - iy)
  const bool xl = t.x0 >= 0 && t.x0 < W, xr = t.x0 + 1 >= 0 && t.x0 + 1 < W;
  const bool yt = t.y0 >= 0 && t.y0 < H, yb = t.y0 + 1 >= 0 && t.y0 + 1 < H;
  const bool in[4] = {xl && yt, xr && yt, xl && yb, xr && yb};
  const float w[4] = {bx * by, ax * by, bx * ay, ax * ay};
  const float dx[4] = {-by, by, -ay, ay}, dy[4] = {-bx, -ax, bx, ax};
  float m = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    t.w[k] = in[k] ? w[k] : 0.f;
    t.dwx[k] = in[k] ? dx[k] : 0.f;
    t.dwy[k] = in[k] ? dy[k] : 0.f;
    if (in[k]) m += w[k];          // grid_sample of the all-ones tensor, same accumulation order
  }
  t.mask = m >= 0.0001f ? 1.f : 0.f;
  return t;
}

constexpr int kChPerThread = 8;

__global__ void __launch_bounds__(256)
warp_fwd_kernel(const float *__restrict__ in, const float *__restrict__ flow, float *__restrict__ out, int B, int C,
                int H, int W, float inv_w1, float inv_h1) {
  const long long HW = (long long)H * W;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long long)B * HW) return;
  const int b = (int)(pix / HW);
  const int r = (int)(pix - (long long)b * HW);
  const int y = r / W, x = r - y * W;
  const Taps t = make_taps(flow[((long long)b * 2) * HW + r], flow[((long long)b * 2 + 1) * HW + r], x, y, H, W,
                           inv_w1, inv_h1);
  // clamped corner offsets: a zero weight makes the value irrelevant, the address must still be legal
  const int xa = min(max(t.x0, 0), W - 1), xb = min(max(t.x0 + 1, 0), W - 1);
  const int ya = min(max(t.y0, 0), H - 1), yb = min(max(t.y0 + 1, 0), H - 1);
  const int o0 = ya * W + xa, o1 = ya * W + xb, o2 = yb * W + xa, o3 = yb * W + xb;
  const int c0 = blockIdx.y * kChPerThread, c1 = min(c0 + kChPerThread, C);
  for (int c = c0; c < c1; ++c) {
    const float *p = in + ((long long)b * C + c) * HW;
    float acc = 0.f;                       // ATen's order: nw, ne, sw, se, each a fused multiply-add
    acc = __fmaf_rn(__ldg(p + o0), t.w[0], acc);
    acc = __fmaf_rn(__ldg(p + o1), t.w[1], acc);
    acc = __fmaf_rn(__ldg(p + o2), t.w[2], acc);
    acc = __fmaf_rn(__ldg(p + o3), t.w[3], acc);
    out[((long long)b * C + c) * HW + r] = acc * t.mask;
  }
}

// grad_in (zero-initialised by the caller of this kernel) += scatter of grad_out * mask * weights;
// grad_flow (zero-initialised) += sum_c grad_out * mask * d(out)/d(ix) * d(ix)/d(u)
__global__ void __launch_bounds__(256)
warp_bwd_kernel(const float *__restrict__ in, const float *__restrict__ flow, const float *__restrict__ gout,
                float *__restrict__ gin, float *__restrict__ gflow, int B, int C, int H, int W, float inv_w1,
                float inv_h1) {
  const long long HW = (long long)H * W;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long long)B * HW) return;
  const int b = (int)(pix / HW);
  const int r = (int)(pix - (long long)b * HW);
  const int y = r / W, x = r - y * W;
  const Taps t = make_taps(flow[((long long)b * 2) * HW + r], flow[((long long)b * 2 + 1) * HW + r], x, y, H, W,
                           inv_w1, inv_h1);
  if (t.mask == 0.f) return;               // masked pixels pass no gradient
  const int xa = min(max(t.x0, 0), W - 1), xb = min(max(t.x0 + 1, 0), W - 1);
  const int ya = min(max(t.y0, 0), H - 1), yb = min(max(t.y0 + 1, 0), H - 1);
  const int o[4] = {ya * W + xa, ya * W + xb, yb * W + xa, yb * W + xb};
  const int c0 = blockIdx.y * kChPerThread, c1 = min(c0 + kChPerThread, C);
  float gx = 0.f, gy = 0.f;
  for (int c = c0; c < c1; ++c) {
    const long long base = ((long long)b * C + c) * HW;
    const float g = gout[base + r];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (t.w[k] != 0.f) atomicAdd(gin + base + o[k], g * t.w[k]);
      const float val = __ldg(in + base + o[k]);
      gx = fmaf(val * t.dwx[k], g, gx);
      gy = fmaf(val * t.dwy[k], g, gy);
    }
  }
  // d(ix)/d(u) = (W / 2) * (2 / (W - 1))   (grid_sample's un-normalisation times PWCNet.py:189's normalisation)
  atomicAdd(gflow + ((long long)b * 2) * HW + r, gx * (0.5f * (float)W) * (2.0f * inv_w1));
  atomicAdd(gflow + ((long long)b * 2 + 1) * HW + r, gy * (0.5f * (float)H) * (2.0f * inv_h1));
}

int check(const char *who, int B, int C, int H, int W) {
  B200_CHECK(B >= 0 && C >= 1 && H >= 1 && W >= 1, "%s: bad sizes", who);
  B200_CHECK((long long)H * W < (1ll << 31), "%s: image too large", who);
  return 0;
}

}  // namespace

extern "C" {

int b200corr_warp_forward(const float *in, const float *flow, float *out, int B, int C, int H, int W, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int e = check("warp_forward", B, C, H, W)) return e;
  if (B == 0) return 0;
  B200_CHECK(in && flow && out, "warp_forward: null pointer");
  const long long npix = (long long)B * H * W;
  dim3 grid((unsigned)((npix + 255) / 256), (unsigned)((C + kChPerThread - 1) / kChPerThread));
  warp_fwd_kernel<<<grid, 256, 0, stream>>>(in, flow, out, B, C, H, W, 1.0f / (float)(W > 1 ? W - 1 : 1),
                                            1.0f / (float)(H > 1 ? H - 1 : 1));
  B200_LAUNCH_OK("warp_fwd_kernel");
  return 0;
}

int b200corr_warp_backward(const float *in, const float *flow, const float *grad_out, float *grad_in, float *grad_flow,
                           int B, int C, int H, int W, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int e = check("warp_backward", B, C, H, W)) return e;
  if (B == 0) return 0;
  B200_CHECK(in && flow && grad_out && grad_in && grad_flow, "warp_backward: null pointer");
  const long long npix = (long long)B * H * W;
  B200_CUDA(cudaMemsetAsync(grad_in, 0, sizeof(float) * npix * C, stream));
  B200_CUDA(cudaMemsetAsync(grad_flow, 0, sizeof(float) * npix * 2, stream));
  dim3 grid((unsigned)((npix + 255) / 256), (unsigned)((C + kChPerThread - 1) / kChPerThread));
  warp_bwd_kernel<<<grid, 256, 0, stream>>>(in, flow, grad_out, grad_in, grad_flow, B, C, H, W,
                                            1.0f / (float)(W > 1 ? W - 1 : 1), 1.0f / (float)(H > 1 ? H - 1 : 1));
  B200_LAUNCH_OK("warp_bwd_kernel");
  return 0;
}

}  // extern "C"
