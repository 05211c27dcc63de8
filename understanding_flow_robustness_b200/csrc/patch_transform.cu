// patch_transform.cu -- the universal-patch placement of the attack loop on the GPU.
//
// Reference: patch_attacks/utils_patch.py:257-358 (`circle_transform`: brightness jitter + clip, `patch * mask`,
// scipy `zoom` order 1, `rotate` order 1, paste at a random location into zero canvases x / xm) followed by
// patch_attacks/main.py:537-542 (`adv = (1 - mask_canvas) * img + patch_canvas`, once per frame) and the clamp to
// the image range.  The reference does this on the host with numpy / scipy once per image pair and folds the
// jitter back into the stored patch; here the transform is a function of the CANONICAL patch:
//
//     q(c,j,i)   = clamp(patch[c,j,i] + bright_n, 0, 1) * mask[j,i]
//     (u, v)     = R(-angle_n) * (x - cx_n, y - cy_n) / scale_n + (p-1)/2          patch coordinate of pixel (x, y)
//     canvas_c   = bilinear(q_c ; u, v)   m = bilinear(mask ; u, v)                zeros outside the patch
//     adv_k[c]   = clamp((1 - m) * img_k[c] + canvas_c, 0, 1)                      k = frame 1, frame 2
//
// i.e. exactly `affine_grid` + `grid_sample(bilinear, zeros, align_corners=True)` of [q, mask] with the matrix
// attack.place() builds, so the gradient lands on the canonical patch (tests compare both directions with that
// torch formulation).  One pass over the images forward; the backward is a GATHER: a thread owns one patch pixel
// of one pair, visits the <= 3x3 image pixels whose bilinear footprint covers it, re-evaluates their clamp
// predicate and sums -- no atomics, deterministic; a second kernel adds the pairs up in index order.
//
// Images, outputs and incoming gradients share one (n, 3, H, W) stride set (contiguous or channels_last).
#include "common.cuh"

namespace {

struct Place {
  float ca, sa;      // cos(angle) / scale, sin(angle) / scale
  float cx, cy, bright, scale, cosa, sina;
};

__device__ __forceinline__ Place load_place(const float *__restrict__ pl, int n) {
  Place P;
  const float s = pl[n * 5 + 0], a = pl[n * 5 + 1];
  P.cx = pl[n * 5 + 2];
  P.cy = pl[n * 5 + 3];
  P.bright = pl[n * 5 + 4];
  float sn, cs;
  sincosf(a, &sn, &cs);
  P.scale = s;
  P.cosa = cs;
  P.sina = sn;
  P.ca = cs / s;
  P.sa = sn / s;
  return P;
}

// patch coordinate of image pixel (x, y)
__device__ __forceinline__ void to_patch(const Place &P, float half, int x, int y, float &u, float &v) {
  const float dx = (float)x - P.cx, dy = (float)y - P.cy;
  u = __fmaf_rn(P.ca, dx, __fmaf_rn(P.sa, dy, half));
  v = __fmaf_rn(-P.sa, dx, __fmaf_rn(P.ca, dy, half));
}

struct Sample {
  float m, c[3];
};

// bilinear taps of [q, mask] at (u, v); zeros outside the p x p patch
__device__ __forceinline__ Sample sample_patch(const float *__restrict__ patch, const float *__restrict__ mask, int p,
                                               float bright, float u, float v) {
  Sample s = {0.f, {0.f, 0.f, 0.f}};
  if (!(u > -1.f && u < (float)p && v > -1.f && v < (float)p)) return s;
  const float fu = floorf(u), fv = floorf(v);
  const int i0 = (int)fu, j0 = (int)fv;
  const float au = u - fu, av = v - fv;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int i = i0 + (t & 1), j = j0 + (t >> 1);
    if (i < 0 || i >= p || j < 0 || j >= p) continue;
    const float w = ((t & 1) ? au : 1.f - au) * ((t >> 1) ? av : 1.f - av);
    const float mk = __ldg(mask + j * p + i);
    const float wm = w * mk;
    s.m += wm;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float q = fminf(fmaxf(__ldg(patch + (c * p + j) * p + i) + bright, 0.f), 1.f);
      s.c[c] = __fmaf_rn(wm, q, s.c[c]);
    }
  }
  return s;
}

__global__ void __launch_bounds__(256)
patch_compose_fwd_kernel(const float *__restrict__ img1, const float *__restrict__ img2,
                         const float *__restrict__ patch, const float *__restrict__ mask,
                         const float *__restrict__ pl, float *__restrict__ adv1, float *__restrict__ adv2, int n, int H,
                         int W, int p, long long sN, long long sC, long long sH, long long sW) {
  const long long HW = (long long)H * W;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long long)n * HW) return;
  const int b = (int)(pix / HW);
  const int r = (int)(pix - (long long)b * HW);
  const int y = r / W, x = r - y * W;
  const Place P = load_place(pl, b);
  float u, v;
  to_patch(P, 0.5f * (float)(p - 1), x, y, u, v);
  const Sample s = sample_patch(patch, mask, p, P.bright, u, v);
  const long long o = b * sN + y * sH + x * sW;
  const float keep = 1.f - s.m;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float a1 = __fmaf_rn(keep, img1[o + c * sC], s.c[c]);
    const float a2 = __fmaf_rn(keep, img2[o + c * sC], s.c[c]);
    adv1[o + c * sC] = fminf(fmaxf(a1, 0.f), 1.f);
    adv2[o + c * sC] = fminf(fmaxf(a2, 0.f), 1.f);
  }
}

// partial[b][c][j][i] = d(sum_k <g_k, adv_k>) / d patch[c,j,i] restricted to pair b
__global__ void __launch_bounds__(128)
patch_compose_bwd_kernel(const float *__restrict__ img1, const float *__restrict__ img2,
                         const float *__restrict__ patch, const float *__restrict__ mask,
                         const float *__restrict__ pl, const float *__restrict__ g1, const float *__restrict__ g2,
                         float *__restrict__ partial, int n, int H, int W, int p, long long sN, long long sC,
                         long long sH, long long sW) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p * p) return;
  const int j = t / p, i = t - j * p;
  float acc[3] = {0.f, 0.f, 0.f};
  const float mk = __ldg(mask + t);
  if (mk != 0.f) {
    const Place P = load_place(pl, b);
    const float half = 0.5f * (float)(p - 1);
    // image position of patch pixel (i, j): inverse of to_patch
    const float uu = (float)i - half, vv = (float)j - half;
    const float xc = P.cx + P.scale * (P.cosa * uu - P.sina * vv);
    const float yc = P.cy + P.scale * (P.sina * uu + P.cosa * vv);
    // |u - i| < 1 and |v - j| < 1 bound the image offset by scale * (|cos| + |sin|) (+ rounding slack)
    const float rad = P.scale * (fabsf(P.cosa) + fabsf(P.sina)) + 1e-3f;
    const int x_lo = max((int)ceilf(xc - rad), 0), x_hi = min((int)floorf(xc + rad), W - 1);
    const int y_lo = max((int)ceilf(yc - rad), 0), y_hi = min((int)floorf(yc + rad), H - 1);
    bool pass[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {       // derivative of clamp(patch + bright, 0, 1): torch passes it on [min, max]
      const float q = __ldg(patch + c * p * p + t) + P.bright;
      pass[c] = q >= 0.f && q <= 1.f;
    }
    for (int y = y_lo; y <= y_hi; ++y) {
      for (int x = x_lo; x <= x_hi; ++x) {
        float u, v;
        to_patch(P, half, x, y, u, v);
        const float wu = 1.f - fabsf(u - (float)i), wv = 1.f - fabsf(v - (float)j);
        if (wu <= 0.f || wv <= 0.f) continue;
        // the weight of tap (i, j) exactly as the forward formed it (au / 1 - au from floor(u))
        const float fu = floorf(u), fv = floorf(v);
        const float au = u - fu, av = v - fv;
        const float w = (((int)fu == i) ? 1.f - au : au) * (((int)fv == j) ? 1.f - av : av);
        const Sample s = sample_patch(patch, mask, p, P.bright, u, v);
        const long long o = b * sN + y * sH + x * sW;
        const float keep = 1.f - s.m;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float a1 = __fmaf_rn(keep, img1[o + c * sC], s.c[c]);
          const float a2 = __fmaf_rn(keep, img2[o + c * sC], s.c[c]);
          float g = 0.f;
          if (a1 >= 0.f && a1 <= 1.f) g += g1[o + c * sC];
          if (a2 >= 0.f && a2 <= 1.f) g += g2[o + c * sC];
          acc[c] = __fmaf_rn(w, g, acc[c]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] = pass[c] ? acc[c] * mk : 0.f;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) partial[((long long)b * 3 + c) * p * p + t] = acc[c];
}

__global__ void __launch_bounds__(256)
patch_reduce_kernel(const float *__restrict__ partial, float *__restrict__ grad_patch, int n, int elems) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= elems) return;
  float s = 0.f;
  for (int b = 0; b < n; ++b) s += partial[(long long)b * elems + t];      // index order: deterministic
  grad_patch[t] = s;
}

}  // namespace

extern "C" {

size_t b200corr_patch_compose_backward_scratch_bytes(int n, int p) {
  return (size_t)(n > 0 ? n : 0) * 3 * p * p * sizeof(float);
}

int b200corr_patch_compose_forward(const float *img1, const float *img2, const float *patch, const float *mask,
                                   const float *placements, float *adv1, float *adv2, int n, int H, int W, int p,
                                   long long sN, long long sC, long long sH, long long sW, void *stream) {
  B200_CHECK(n >= 0 && H > 0 && W > 0 && p >= 2, "patch_compose_forward: bad sizes n=%d H=%d W=%d p=%d", n, H, W, p);
  if (n == 0) return 0;
  B200_CHECK(img1 && img2 && patch && mask && placements && adv1 && adv2, "patch_compose_forward: null pointer");
  const long long total = (long long)n * H * W;
  patch_compose_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      img1, img2, patch, mask, placements, adv1, adv2, n, H, W, p, sN, sC, sH, sW);
  B200_LAUNCH_OK("patch_compose_fwd_kernel");
  return 0;
}

int b200corr_patch_compose_backward(const float *img1, const float *img2, const float *patch, const float *mask,
                                    const float *placements, const float *grad_adv1, const float *grad_adv2,
                                    float *grad_patch, float *scratch, size_t scratch_bytes, int n, int H, int W, int p,
                                    long long sN, long long sC, long long sH, long long sW, void *stream) {
  B200_CHECK(n >= 0 && H > 0 && W > 0 && p >= 2, "patch_compose_backward: bad sizes n=%d H=%d W=%d p=%d", n, H, W, p);
  B200_CHECK(grad_patch, "patch_compose_backward: null grad_patch");
  const int elems = 3 * p * p;
  if (n == 0) {
    B200_CUDA(cudaMemsetAsync(grad_patch, 0, elems * sizeof(float), (cudaStream_t)stream));
    return 0;
  }
  B200_CHECK(img1 && img2 && patch && mask && placements && grad_adv1 && grad_adv2, "patch_compose_backward: null pointer");
  B200_CHECK(scratch && scratch_bytes >= b200corr_patch_compose_backward_scratch_bytes(n, p),
             "patch_compose_backward: scratch too small (%zu bytes)", scratch_bytes);
  dim3 grid((unsigned)b200::ceil_div(p * p, 128), (unsigned)n);
  patch_compose_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(img1, img2, patch, mask, placements, grad_adv1,
                                                                   grad_adv2, scratch, n, H, W, p, sN, sC, sH, sW);
  B200_LAUNCH_OK("patch_compose_bwd_kernel");
  patch_reduce_kernel<<<b200::ceil_div(elems, 256), 256, 0, (cudaStream_t)stream>>>(scratch, grad_patch, n, elems);
  B200_LAUNCH_OK("patch_reduce_kernel");
  return 0;
}

}  // extern "C"
