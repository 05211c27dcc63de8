// raft_lookup.cu -- RAFT correlation lookup (forward + backward) and pyramid backward.
//
// Replaces CorrBlock.__call__ + bilinear_sampler of the reference (models/raft/corr.py:72-96,
// models/raft/utils/utils.py:62-76): per pyramid level a CPU-built (2r+1)^2 offset grid copied to
// the device, one F.grid_sample launch, then cat + permute + contiguous -- 4 grid_sample launches,
// 4 H2D copies and 2 extra passes over the (B, 324, H, W) result per RAFT iteration.
//
// Here: ONE launch per lookup for all levels.  A CTA owns 32 consecutive query pixels of one level:
//   phase 1: the (2r+4)^2 neighbourhood of every query is staged in shared memory (rows of the
//            query's own H_l x W_l slice; zero outside the slice = grid_sample's zero padding);
//   phase 2: warp = tap subset, lane = query, so every store of out[b, l*81 + k, q] is a full
//            128-byte line; the result is written once, in its final (B, L*(2r+1)^2, H, W) layout.
// Channel order k = i*(2r+1) + j with i the x offset and j the y offset (corr.py:80-86).
//
// Coordinate arithmetic, mode B200CORR_LOOKUP_GRIDSAMPLE: the reference normalises the pixel
// coordinate to [-1,1] (utils.py:65-67) and grid_sample un-normalises it again
// (ATen grid_sampler_unnormalize, align_corners=True: ((g + 1) / 2) * (size - 1)); both steps are
// reproduced operation by operation in fp32 (no FMA contraction) so that integer coordinates come
// back as e.g. 79.99999 exactly as in the reference.  Mode B200CORR_LOOKUP_DIRECT samples at the
// pixel coordinate itself (what alt_cuda_corr does).
//
// Backward (what autograd derives for the reference): the bilinear weights
// are scattered into the query's own slice of a dense per-level gradient volume.  Each slice is
// touched by exactly one CTA per launch, so the scatter is first reduced in shared memory and then
// added with plain coalesced read-modify-writes -- no global atomics.  Coordinates get no gradient
// (raft.py:188 detaches them).
#include "common.cuh"

namespace {

constexpr int kMaxLevels = 8;
constexpr int QT = 32;  // queries per CTA

struct LookupParams {
  const float *lvl[kMaxLevels];
  float *glvl[kMaxLevels];
  int LH[kMaxLevels], LW[kMaxLevels];
  int num_levels, B, HW, radius, mode;
};

// pixel coordinate the reference ends up sampling at, along one axis of size `size`
__device__ __forceinline__ float sample_coord(float c, int lvl, int off, int size, int mode) {
  // corr.py:84-86: centroid / 2**i  + delta   (division by a power of two is exact)
  const float x = __fadd_rn(__fmul_rn(c, 1.0f / (float)(1 << lvl)), (float)off);
  if (mode == B200CORR_LOOKUP_DIRECT) return x;
  const float sm1 = (float)(size - 1);
  // utils.py:66: 2 * x / (W - 1) - 1
  const float g = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, x), sm1), 1.0f);
  // grid_sampler_unnormalize, align_corners=True
  return __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.0f), 2.0f), sm1);
}

template <int R>
struct Geo {
  static constexpr int N = 2 * R + 1;      // taps per axis
  static constexpr int WS = 2 * R + 4;     // staged window per axis
  static constexpr int WSTRIDE = WS * WS + 1;
  // vector path (level width % 4 == 0): a window row is fetched as NV4 aligned float4 covering
  // [ox & ~3, ...) and stored shifted by (ox & 3); odd row / query pitches keep both the staging
  // stores (lane = window row) and the sampling loads (lane = query) bank-conflict free
  static constexpr int NV4 = (WS + 3 + 3) / 4;
  static constexpr int VROW = WS | 1;
  static constexpr int VSTRIDE = (WS * VROW) | 1;
};

// per-query per-axis tap tables in shared memory
template <int R>
struct TapTables {
  int x0[QT][Geo<R>::N];   // floor of the sample coordinate, relative to the window origin
  float ax[QT][Geo<R>::N]; // fractional part
  int y0[QT][Geo<R>::N];
  float ay[QT][Geo<R>::N];
  int ox[QT], oy[QT];      // window origin in the level's pixel coordinates
};

template <int R>
__device__ __forceinline__ void build_taps(TapTables<R> &tt, const float *coords, int b, int q0,
                                           int HW, int lvl, int LH, int LW, int mode) {
  constexpr int N = Geo<R>::N;
  // one task per (query, axis, tap): 32 * 2 * N tasks spread over the whole CTA (the two IEEE
  // divisions of the grid_sample round trip make a tap ~80 instructions)
  for (int i = threadIdx.x; i < QT * 2 * N; i += blockDim.x) {
    const int t = i % N;
    const int qa = i / N;
    const int qi = qa >> 1, axis = qa & 1;
    const int q = q0 + qi;
    float c = 0.f;
    if (q < HW) c = coords[((size_t)b * 2 + axis) * HW + q];
    const int size = axis == 0 ? LW : LH;
    // window origin from the un-rounded centre: floor(c / 2^l) - R - 1
    const float cl = c * (1.0f / (float)(1 << lvl));
    float fo = floorf(cl);
    if (!(fabsf(fo) < 1e8f)) fo = -1e8f;  // non-finite / absurd coordinates: everything out of range
    const int org = (int)fo - R - 1;
    if (t == 0) {
      if (axis == 0) tt.ox[qi] = org; else tt.oy[qi] = org;
    }
    const float x = sample_coord(c, lvl, t - R, size, mode);
    float fx = floorf(x);
    int rel;
    float frac;
    if (fabsf(fx) < 1e8f) {
      rel = (int)fx - org;
      frac = x - fx;
      // the round trip moves x by a few ulp at most: rel is within [0, WS-2]; clamp defensively
      if (rel < 0 || rel > Geo<R>::WS - 2) { rel = -1; frac = 0.f; }
    } else {
      rel = 0; frac = x - x;  // NaN propagates like in the reference
    }
    if (axis == 0) { tt.x0[qi][t] = rel; tt.ax[qi][t] = frac; }
    else           { tt.y0[qi][t] = rel; tt.ay[qi][t] = frac; }
  }
}

template <int R, bool VEC>
__global__ void __launch_bounds__(256)
lookup_fwd_kernel(const LookupParams p, const float *__restrict__ coords, float *__restrict__ out) {
  constexpr int N = Geo<R>::N, WS = Geo<R>::WS;
  constexpr int ROW = VEC ? Geo<R>::VROW : WS;
  constexpr int QSTRIDE = VEC ? Geo<R>::VSTRIDE : Geo<R>::WSTRIDE;
  __shared__ float win[QT * QSTRIDE];
  __shared__ TapTables<R> tt;
  const int lvl = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * QT;
  const int LH = p.LH[lvl], LW = p.LW[lvl];
  const float *vol = p.lvl[lvl];

  build_taps<R>(tt, coords, b, q0, p.HW, lvl, LH, LW, p.mode);
  __syncthreads();

  // phase 1: stage windows, thread = (query, window row)
  for (int i = threadIdx.x; i < QT * WS; i += blockDim.x) {
    const int qi = i / WS, r = i - qi * WS;
    const int q = q0 + qi;
    const int y = tt.oy[qi] + r, xo = tt.ox[qi];
    const bool row_ok = q < p.HW && y >= 0 && y < LH;
    if (VEC) {
      float *dst = win + qi * QSTRIDE + r * ROW;
      const int a0 = xo & ~3;  // aligned first column (floor to a multiple of 4, also for negatives)
      const float4 *src = reinterpret_cast<const float4 *>(
          vol + (((size_t)b * p.HW + (row_ok ? q : 0)) * LH + (row_ok ? y : 0)) * LW);
      float v[Geo<R>::NV4 * 4];
#pragma unroll
      for (int k = 0; k < Geo<R>::NV4; ++k) {
        const int x = a0 + 4 * k;
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row_ok && x >= 0 && x < LW) val = __ldg(src + (x >> 2));
        v[4 * k] = val.x; v[4 * k + 1] = val.y; v[4 * k + 2] = val.z; v[4 * k + 3] = val.w;
      }
      switch (xo & 3) {  // compile-time register indices inside each case
        case 0:
#pragma unroll
          for (int c = 0; c < WS; ++c) dst[c] = v[c];
          break;
        case 1:
#pragma unroll
          for (int c = 0; c < WS; ++c) dst[c] = v[c + 1];
          break;
        case 2:
#pragma unroll
          for (int c = 0; c < WS; ++c) dst[c] = v[c + 2];
          break;
        default:
#pragma unroll
          for (int c = 0; c < WS; ++c) dst[c] = v[c + 3];
          break;
      }
    } else {
      float *dst = win + qi * QSTRIDE + r * ROW;
      if (row_ok) {
        const float *src = vol + (((size_t)b * p.HW + q) * LH + y) * LW;
#pragma unroll
        for (int c = 0; c < WS; ++c) {
          const int x = xo + c;
          dst[c] = (x >= 0 && x < LW) ? src[x] : 0.f;
        }
      } else {
#pragma unroll
        for (int c = 0; c < WS; ++c) dst[c] = 0.f;
      }
    }
  }
  __syncthreads();

  // phase 2: lane = query, warp = y-offset row j; the x tap tables of the query live in registers
  // and the two window rows are walked left to right, reusing the previous column pair whenever the
  // next tap starts one pixel further (always, up to the reference's coordinate round-trip noise)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = q0 + lane;
  const float *w = win + lane * QSTRIDE;
  const int nchan = p.num_levels * N * N;
  int rxs[N];
  float axs[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    rxs[i] = tt.x0[lane][i];
    axs[i] = tt.ax[lane][i];
  }
  for (int j = warp; j < N; j += 8) {
    const int ry = tt.y0[lane][j];
    const float ay = tt.ay[lane][j], by = 1.f - ay;
    const float *r0 = w + (ry >= 0 ? ry : 0) * ROW;
    int prx = -100;
    float c00 = 0.f, c01 = 0.f, c10 = 0.f, c11 = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const int rx = rxs[i];
      float v = 0.f;
      if (rx >= 0 && ry >= 0) {
        if (rx == prx + 1) {
          c00 = c01;
          c10 = c11;
        } else {
          c00 = r0[rx];
          c10 = r0[ROW + rx];
        }
        c01 = r0[rx + 1];
        c11 = r0[ROW + rx + 1];
        prx = rx;
        // grid_sample: nw = (x1-x)(y1-y), ne = (x-x0)(y1-y), sw = (x1-x)(y-y0), se = (x-x0)(y-y0)
        const float ax = axs[i], bx = 1.f - ax;
        v = c00 * (bx * by);
        v += c01 * (ax * by);
        v += c10 * (bx * ay);
        v += c11 * (ax * ay);
      }
      if (q < p.HW) out[((size_t)b * nchan + lvl * N * N + i * N + j) * p.HW + q] = v;
    }
  }
}

template <int R>
__global__ void __launch_bounds__(256)
lookup_bwd_kernel(const LookupParams p, const float *__restrict__ coords,
                  const float *__restrict__ gout) {
  constexpr int N = Geo<R>::N, WS = Geo<R>::WS, WSTRIDE = Geo<R>::WSTRIDE;
  __shared__ float win[QT * WSTRIDE];
  __shared__ TapTables<R> tt;
  const int lvl = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * QT;
  const int LH = p.LH[lvl], LW = p.LW[lvl];
  float *gvol = p.glvl[lvl];

  build_taps<R>(tt, coords, b, q0, p.HW, lvl, LH, LW, p.mode);
  for (int i = threadIdx.x; i < QT * WSTRIDE; i += blockDim.x) win[i] = 0.f;
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = q0 + lane;
  float *w = win + lane * WSTRIDE;
  const int nchan = p.num_levels * N * N;
  // lane = query: two taps of the same query are handled by different warps -> shared atomics
  for (int k = warp; k < N * N; k += 8) {
    const int i = k / N, j = k - i * N;
    const int rx = tt.x0[lane][i], ry = tt.y0[lane][j];
    if (q < p.HW && rx >= 0 && ry >= 0) {
      const float g = gout[((size_t)b * nchan + lvl * N * N + k) * p.HW + q];
      const float ax = tt.ax[lane][i], ay = tt.ay[lane][j];
      const float bx = 1.f - ax, by = 1.f - ay;
      float *c = w + ry * WS + rx;
      atomicAdd(c, g * (bx * by));
      atomicAdd(c + 1, g * (ax * by));
      atomicAdd(c + WS, g * (bx * ay));
      atomicAdd(c + WS + 1, g * (ax * ay));
    }
  }
  __syncthreads();
  // add the window into the query's slice (only this CTA touches these slices in this launch)
  for (int i = threadIdx.x; i < QT * WS; i += blockDim.x) {
    const int qi = i / WS, r = i - qi * WS;
    const int qq = q0 + qi;
    const int y = tt.oy[qi] + r, xo = tt.ox[qi];
    if (qq < p.HW && y >= 0 && y < LH) {
      float *dst = gvol + (((size_t)b * p.HW + qq) * LH + y) * LW;
      const float *src = win + qi * WSTRIDE + r * WS;
#pragma unroll
      for (int c = 0; c < WS; ++c) {
        const int x = xo + c;
        if (x >= 0 && x < LW && src[c] != 0.f) dst[x] += src[c];
      }
    }
  }
}

// fine[q, y, x] += coarse[q, y/2, x/2] / 4   for y < 2*Hc, x < 2*Wc   (backward of avg_pool2d(2,2))
__global__ void __launch_bounds__(256)
pool_bwd_kernel(float *__restrict__ fine, const float *__restrict__ coarse, long long Q, int Hf, int Wf) {
  const int Hc = Hf / 2, Wc = Wf / 2;
  const long long total = Q * Hf * Wf;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wf);
    const long long t = i / Wf;
    const int y = (int)(t % Hf);
    const long long q = t / Hf;
    if (y < 2 * Hc && x < 2 * Wc) fine[i] += 0.25f * coarse[(q * Hc + (y >> 1)) * Wc + (x >> 1)];
  }
}

int fill_params(LookupParams &p, const float *const *lv, float *const *glv, int num_levels, int B,
                int H, int W, int radius, int mode, const char *who) {
  B200_CHECK(num_levels >= 1 && num_levels <= kMaxLevels, "%s: num_levels must be in [1, %d]", who,
             kMaxLevels);
  B200_CHECK(radius >= 1 && radius <= 4, "%s: radius %d not instantiated (1..4)", who, radius);
  B200_CHECK(mode == B200CORR_LOOKUP_GRIDSAMPLE || mode == B200CORR_LOOKUP_DIRECT, "%s: bad mode", who);
  B200_CHECK(B >= 0 && H >= 1 && W >= 1, "%s: bad sizes", who);
  p.num_levels = num_levels; p.B = B; p.HW = H * W; p.radius = radius; p.mode = mode;
  int h = H, w = W;
  for (int l = 0; l < kMaxLevels; ++l) {
    p.lvl[l] = nullptr; p.glvl[l] = nullptr; p.LH[l] = 0; p.LW[l] = 0;
    if (l < num_levels) {
      p.LH[l] = h; p.LW[l] = w;
      B200_CHECK(h >= 1 && w >= 1, "%s: pyramid level %d is empty (%dx%d input)", who, l, H, W);
      if (lv) { B200_CHECK(lv[l], "%s: null level %d", who, l); p.lvl[l] = lv[l]; }
      if (glv) { B200_CHECK(glv[l], "%s: null gradient level %d", who, l); p.glvl[l] = glv[l]; }
      h /= 2; w /= 2;
    }
  }
  return 0;
}

}  // namespace

extern "C" {

int b200corr_lookup_forward(const float *const *h_levels, int num_levels, const float *coords,
                            float *out, int B, int H, int W, int radius, int mode, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LookupParams p;
  if (int e = fill_params(p, h_levels, nullptr, num_levels, B, H, W, radius, mode, "lookup_forward")) return e;
  if (B == 0) return 0;
  B200_CHECK(coords && out, "lookup_forward: null pointer");
  dim3 grid((p.HW + QT - 1) / QT, num_levels, B);
  bool vec = true;  // aligned float4 staging needs every level width % 4 == 0 and 16-byte aligned bases
  for (int l = 0; l < num_levels; ++l) vec = vec && p.LW[l] % 4 == 0 && ((uintptr_t)p.lvl[l] & 15) == 0;
#define LOOKUP_FWD(RR)                                                              \
  if (vec) lookup_fwd_kernel<RR, true><<<grid, 256, 0, stream>>>(p, coords, out);   \
  else lookup_fwd_kernel<RR, false><<<grid, 256, 0, stream>>>(p, coords, out)
  switch (radius) {
    case 1: LOOKUP_FWD(1); break;
    case 2: LOOKUP_FWD(2); break;
    case 3: LOOKUP_FWD(3); break;
    default: LOOKUP_FWD(4); break;
  }
#undef LOOKUP_FWD
  B200_LAUNCH_OK("lookup_fwd_kernel");
  return 0;
}

int b200corr_lookup_backward(float *const *h_grad_levels, int num_levels, const float *coords,
                             const float *grad_out, int B, int H, int W, int radius, int mode,
                             void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LookupParams p;
  if (int e = fill_params(p, nullptr, h_grad_levels, num_levels, B, H, W, radius, mode, "lookup_backward")) return e;
  if (B == 0) return 0;
  B200_CHECK(coords && grad_out, "lookup_backward: null pointer");
  dim3 grid((p.HW + QT - 1) / QT, num_levels, B);
  switch (radius) {
    case 1: lookup_bwd_kernel<1><<<grid, 256, 0, stream>>>(p, coords, grad_out); break;
    case 2: lookup_bwd_kernel<2><<<grid, 256, 0, stream>>>(p, coords, grad_out); break;
    case 3: lookup_bwd_kernel<3><<<grid, 256, 0, stream>>>(p, coords, grad_out); break;
    default: lookup_bwd_kernel<4><<<grid, 256, 0, stream>>>(p, coords, grad_out); break;
  }
  B200_LAUNCH_OK("lookup_bwd_kernel");
  return 0;
}

int b200corr_pyramid_backward(float *const *h_grad_levels, int num_levels, int B, int H, int W,
                              void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  LookupParams p;
  if (int e = fill_params(p, nullptr, h_grad_levels, num_levels, B, H, W, 1, 0, "pyramid_backward")) return e;
  if (B == 0) return 0;
  const long long Q = (long long)B * H * W;
  for (int l = num_levels - 1; l >= 1; --l) {
    const long long total = Q * p.LH[l - 1] * p.LW[l - 1];
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)b200::num_sms() * 16;
    pool_bwd_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, stream>>>(p.glvl[l - 1], p.glvl[l], Q,
                                                                          p.LH[l - 1], p.LW[l - 1]);
    B200_LAUNCH_OK("pool_bwd_kernel");
  }
  return 0;
}

}  // extern "C"
