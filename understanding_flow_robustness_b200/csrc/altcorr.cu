// altcorr.cu -- on-the-fly windowed correlation: drop-in for alt_cuda_corr.forward / .backward.
//
// Reference: models/alt_cuda_corr/correlation_kernel.cu:18-119 (forward), :122-256 (backward),
// checks and bindings models/alt_cuda_corr/correlation.cpp:23-54.  Semantics:
//   s[iy][ix] = <fmap1[b,h,w,:], fmap2[b, floor(y)-r+iy, floor(x)-r+ix, :]>   (0 outside fmap2)
//   corr[b,n, ix*(2r+1)+iy, h,w] = (1-dy)(1-dx) s[iy][ix] + (1-dy)dx s[iy][ix+1]
//                                 + dy(1-dx) s[iy+1][ix] + dy dx s[iy+1][ix+1]
// with (x, y) = coords[b,n,h,w,:], dx = x - floor(x), dy = y - floor(y); no 1/sqrt(C) factor.
//
// The reference stages 32-channel tiles of both maps per 4x8 query block and loops the (2r+2)^2
// window serially with one barrier per window point.  Here one warp owns one query: the query's
// feature vector lives in registers (lane = channel slice), every window pixel of fmap2 is one
// fully coalesced C*4-byte read (NHWC), all reads of a window row in flight together; the partial dot
// products of the row are reduced across lanes by a shuffle reduce-scatter; the 8 queries of a CTA then
// write their (2r+1)^2 outputs as 32-byte sectors.
#include "common.cuh"

namespace {

constexpr int kMaxVec = 4;   // float4 per lane: C <= 512

template <int R>
struct AGeo {
  static constexpr int RD = 2 * R + 1;
  static constexpr int WN = 2 * R + 2;
  static constexpr int NW = WN * WN;
  static constexpr int NCHUNK = (NW + 31) / 32;
};

template <int R, int NV>
__global__ void __launch_bounds__(256)
altcorr_fwd_kernel(const float *__restrict__ f1, const float *__restrict__ f2,
                   const float *__restrict__ coords, float *__restrict__ corr, int B, int N, int H1,
                   int W1, int H2, int W2, int C) {
  using G = AGeo<R>;
  __shared__ float red[8][G::WN][33];      // per-warp [window column][lane] partial sums of ONE window row
  __shared__ float S[8][G::NCHUNK * 32];
  __shared__ float O[G::RD * G::RD][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HW1 = H1 * W1;
  const int q0 = blockIdx.x * 8, n = blockIdx.y, b = blockIdx.z;
  const int q = q0 + warp;
  const bool q_ok = q < HW1;
  const int nvec = C >> 2;

  float x = 0.f, y = 0.f;
  float4 a[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (q_ok) {
    const float *cp = coords + (((size_t)b * N + n) * HW1 + q) * 2;
    x = cp[0];
    y = cp[1];
    const float4 *p1 = reinterpret_cast<const float4 *>(f1 + ((size_t)b * HW1 + q) * C);
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (lane + 32 * i < nvec) a[i] = p1[lane + 32 * i];
  }
  const int rt_zero = B >> 30;   // 0 for every real batch size; the compiler cannot know
  float fxf = floorf(x), fyf = floorf(y);
  const float dx = x - fxf, dy = y - fyf;
  if (!(fabsf(fxf) < 1e8f)) fxf = -1e8f;
  if (!(fabsf(fyf) < 1e8f)) fyf = -1e8f;
  const int fx = (int)fxf, fy = (int)fyf;

  // One window row at a time (a real loop: the fully unrolled 100-pixel version thrashed the instruction cache).
  // What bounded this kernel until late in round 2 was not L2 bandwidth (as rounds 1-2 read the 8.3 TB/s of L2 -> L1
  // traffic; scripts/probes/l2_read_probe.cu: plain loads pull 16 TB/s out of the L2) but a chain of 100 L2 round
  // trips per query: the compiler gave every window pixel its own branch region -- 2 loads, then the 8 FMAs that wait
  // for them.  Now all loads of a row are issued, predicated, before the first FMA (inline PTX + a data-dependent
  // scheduling fence, below), and the row's WN partial sums are reduced across the lanes through a per-warp tile read
  // back by 30 lanes (three per column) instead of 10: 0.22 -> 0.18 ms per
  // level at B=4 (10 round trips per query, 16 warps per SM at 117 registers); interior batches (warp-uniform test) use
  // plain loads off one base pointer instead of zero-init + predicate per load, and the dot products run as packed
  // fp32x2 FMAs: 0.18 -> 0.148 ms.  Also measured: the next row's loads issued before this row's reduction (156
  // registers, or 128 with spills): 0.194 ms; half-row batches for 3 CTAs per SM: 0.158; 256-bit loads (two adjacent
  // vectors per lane): 0.166; a shuffle reduce-scatter instead of the tile: 0.148 (same); round 2's earlier attempts
  // (__syncthreads per row for L1 sharing 0.221 vs 0.212; CTA-wide staging of the union window box 0.31 ms).
  // pixels whose loads are in flight together: the whole row while that is <= 80 registers (C <= 256), else half
#ifndef B200_ALT_PB_FULL_NV
#define B200_ALT_PB_FULL_NV 2
#endif
  constexpr int PB = NV <= B200_ALT_PB_FULL_NV ? G::WN : G::WN / 2;
  static_assert(G::WN % PB == 0, "batches tile the window row");
#pragma unroll 1
  for (int iy = 0; iy < G::WN; ++iy) {
    const int y2 = fy - R + iy;
    const bool row_ok = q_ok && y2 >= 0 && y2 < H2;
    const float4 *prow = reinterpret_cast<const float4 *>(f2 + (((size_t)b * H2 + (row_ok ? y2 : 0)) * W2) * C);
    float acc[G::WN];
#pragma unroll
    for (int x0 = 0; x0 < G::WN; x0 += PB) {
      float4 v[PB][NV];
      const int xl = fx - R + x0;
      if (row_ok && xl >= 0 && xl + PB <= W2 && nvec == 32 * NV) {
        // interior batch, every lane loaded (warp-uniform test): plain loads off one base pointer
        const float4 *p2 = prow + (size_t)xl * nvec + lane;
#pragma unroll
        for (int px = 0; px < PB; ++px)
#pragma unroll
          for (int i = 0; i < NV; ++i)
            asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(v[px][i].x), "=f"(v[px][i].y), "=f"(v[px][i].z), "=f"(v[px][i].w)
                         : "l"(p2 + px * nvec + 32 * i));
      } else {
#pragma unroll
      for (int px = 0; px < PB; ++px) {
        const int x2 = xl + px;
        const bool ok = row_ok && x2 >= 0 && x2 < W2;
        const float4 *p2 = prow + (size_t)(ok ? x2 : 0) * nvec;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int pr = (ok && lane + 32 * i < nvec) ? 1 : 0;
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t"
              "mov.f32 %0, 0f00000000;\n\tmov.f32 %1, 0f00000000;\n\tmov.f32 %2, 0f00000000;\n\tmov.f32 %3, 0f00000000;\n\t"
              "@p ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
              : "=f"(v[px][i].x), "=f"(v[px][i].y), "=f"(v[px][i].z), "=f"(v[px][i].w)
              : "l"(p2 + lane + 32 * i), "r"(pr));
        }
      }
      }
      // scheduling fence: every dot product starts from a zero that is COMPUTED from all of the batch's loads (their
      // bits or-ed, & a run-time zero), so ptxas cannot sink the first FMAs in between the loads -- in-order issue
      // would stall on them with most of the row still unrequested (it did: 4-6 loads in flight instead of 20)
      int bits = 0;
#pragma unroll
      for (int px = 0; px < PB; ++px)
#pragma unroll
        for (int i = 0; i < NV; ++i) bits |= __float_as_int(v[px][i].w);
      const float dep = __int_as_float(bits & rt_zero);
#pragma unroll
      for (int px = 0; px < PB; ++px) {
        // packed fp32x2 FMAs (sm_100): even and odd channels accumulate side by side, one add joins them
        float2 s2 = make_float2(dep, 0.f);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          s2 = __ffma2_rn(make_float2(a[i].x, a[i].y), make_float2(v[px][i].x, v[px][i].y), s2);
          s2 = __ffma2_rn(make_float2(a[i].z, a[i].w), make_float2(v[px][i].z, v[px][i].w), s2);
        }
        acc[x0 + px] = s2.x + s2.y;
      }
    }
    // cross-lane reduction of the row's WN partial sums through a per-warp tile [column][lane] (pitch 33): three lanes
    // per column add 11 + 11 + 10 of the 32 partials each (bank = column + 11 m + t: conflict-free), two shuffles join
    // them -- 36 instructions per row with every lane busy (a shuffle reduce-scatter: 62; ten lanes adding 32 each: 74)
    __syncwarp();   // the previous row's readers are done with the tile
#pragma unroll
    for (int ix = 0; ix < G::WN; ++ix) red[warp][ix][lane] = acc[ix];
    __syncwarp();
    {
      const int col = lane / 3, m = lane - 3 * col;
      float part = 0.f;
      if (col < G::WN) {
        const float *t = &red[warp][col][11 * m];
#pragma unroll
        for (int k = 0; k < 10; ++k) part += t[k];
        if (m < 2) part += t[10];
      }
      const float p1 = __shfl_down_sync(0xffffffffu, part, 1), p2 = __shfl_down_sync(0xffffffffu, part, 2);
      if (m == 0 && col < G::WN) S[warp][iy * G::WN + col] = (part + p1) + p2;
    }
  }
  __syncwarp();
  // bilinear combination -> (2r+1)^2 outputs, channel = ix * RD + iy
  for (int k = lane; k < G::RD * G::RD; k += 32) {
    const int ix = k / G::RD, iy = k - ix * G::RD;
    const float *s = S[warp];
    const float v = (1.f - dy) * (1.f - dx) * s[iy * G::WN + ix] + (1.f - dy) * dx * s[iy * G::WN + ix + 1] +
                    dy * (1.f - dx) * s[(iy + 1) * G::WN + ix] + dy * dx * s[(iy + 1) * G::WN + ix + 1];
    O[k][warp] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < G::RD * G::RD * 8; i += 256) {
    const int k = i >> 3, qi = i & 7;
    if (q0 + qi < HW1) corr[(((size_t)b * N + n) * (G::RD * G::RD) + k) * HW1 + q0 + qi] = O[k][qi];
  }
}

template <int R, int NV>
__global__ void __launch_bounds__(256)
altcorr_bwd_kernel(const float *__restrict__ f1, const float *__restrict__ f2,
                   const float *__restrict__ coords, const float *__restrict__ cgrad,
                   float *__restrict__ f1g, float *__restrict__ f2g, int B, int N, int H1, int W1,
                   int H2, int W2, int C) {
  using G = AGeo<R>;
  __shared__ float GS[8][G::NCHUNK * 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HW1 = H1 * W1;
  const int q = blockIdx.x * 8 + warp, b = blockIdx.z;
  if (q >= HW1) return;
  const int nvec = C >> 2;
  float4 a[NV], ga[NV];
  const float4 *p1 = reinterpret_cast<const float4 *>(f1 + ((size_t)b * HW1 + q) * C);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    ga[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane + 32 * i < nvec) a[i] = p1[lane + 32 * i];
  }
  for (int n = 0; n < N; ++n) {
    const float *cp = coords + (((size_t)b * N + n) * HW1 + q) * 2;
    const float x = cp[0], y = cp[1];
    float fxf = floorf(x), fyf = floorf(y);
    const float dx = x - fxf, dy = y - fyf;
    if (!(fabsf(fxf) < 1e8f)) fxf = -1e8f;
    if (!(fabsf(fyf) < 1e8f)) fyf = -1e8f;
    const int fx = (int)fxf, fy = (int)fyf;
    // gradient w.r.t. the window dot products (gather form of correlation_kernel.cu:203-218)
    const float *gq = cgrad + (((size_t)b * N + n) * (G::RD * G::RD)) * HW1 + q;
    __syncwarp();
    for (int idx = lane; idx < G::NW; idx += 32) {
      const int iy = idx / G::WN, ix = idx - iy * G::WN;
      float g = 0.f;
      if (ix < G::RD && iy < G::RD) g += (1.f - dy) * (1.f - dx) * gq[(size_t)(ix * G::RD + iy) * HW1];
      if (ix >= 1 && iy < G::RD) g += (1.f - dy) * dx * gq[(size_t)((ix - 1) * G::RD + iy) * HW1];
      if (ix < G::RD && iy >= 1) g += dy * (1.f - dx) * gq[(size_t)(ix * G::RD + iy - 1) * HW1];
      if (ix >= 1 && iy >= 1) g += dy * dx * gq[(size_t)((ix - 1) * G::RD + iy - 1) * HW1];
      GS[warp][idx] = g;
    }
    __syncwarp();
#pragma unroll 1
    for (int iy = 0; iy < G::WN; ++iy) {
      const int y2 = fy - R + iy;
      if (y2 < 0 || y2 >= H2) continue;        // warp-uniform
      const size_t rowoff = (((size_t)b * H2 + y2) * W2) * C;
      // all loads of the window row first (clamped addresses, 2*WN in flight per lane), then the FMAs and the
      // vector reductions
      float4 v[G::WN][NV];
#pragma unroll
      for (int ix = 0; ix < G::WN; ++ix) {
        const int x2 = min(max(fx - R + ix, 0), W2 - 1);
        const float4 *p2 = reinterpret_cast<const float4 *>(f2 + rowoff + (size_t)x2 * C);
#pragma unroll
        for (int i = 0; i < NV; ++i) v[ix][i] = __ldg(p2 + min(lane + 32 * i, nvec - 1));
      }
#pragma unroll
      for (int ix = 0; ix < G::WN; ++ix) {
        const int x2 = fx - R + ix;
        const float g = GS[warp][iy * G::WN + ix];
        if (g == 0.f || x2 < 0 || x2 >= W2) continue;
        float *g2 = f2g + rowoff + (size_t)x2 * C;
#pragma unroll
        for (int i = 0; i < NV; ++i)
          if (lane + 32 * i < nvec) {
            ga[i].x = fmaf(g, v[ix][i].x, ga[i].x);
            ga[i].y = fmaf(g, v[ix][i].y, ga[i].y);
            ga[i].z = fmaf(g, v[ix][i].z, ga[i].z);
            ga[i].w = fmaf(g, v[ix][i].w, ga[i].w);
            // one 16-byte vector reduction per lane (sm_90+) instead of four scalar atomics
            float *d = g2 + 4 * (lane + 32 * i);
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(g * a[i].x), "f"(g * a[i].y),
                         "f"(g * a[i].z), "f"(g * a[i].w)
                         : "memory");
          }
      }
    }
  }
  float4 *o = reinterpret_cast<float4 *>(f1g + ((size_t)b * HW1 + q) * C);
#pragma unroll
  for (int i = 0; i < NV; ++i)
    if (lane + 32 * i < nvec) o[lane + 32 * i] = ga[i];
}

int check_alt(const char *who, int B, int N, int H1, int W1, int H2, int W2, int C, int radius) {
  B200_CHECK(B >= 0 && B < (1 << 30) && N >= 1 && H1 >= 1 && W1 >= 1 && H2 >= 1 && W2 >= 1, "%s: bad sizes", who);   // B >> 30 is the forward kernel's run-time zero
  B200_CHECK(C >= 4 && C % 4 == 0 && C <= 128 * kMaxVec, "%s: C must be a multiple of 4, <= %d", who,
             128 * kMaxVec);
  B200_CHECK(radius >= 1 && radius <= 4, "%s: radius %d not instantiated (1..4)", who, radius);
  return 0;
}

#define ALT_DISPATCH(KERNEL, SMEM, ...)                                                         \
  do {                                                                                      \
    const int nv = (C / 4 + 31) / 32;                                                       \
    const size_t smem = (SMEM);                                                             \
    switch (radius * 8 + nv) {                                                              \
      case 1 * 8 + 1: KERNEL<1, 1><<<grid, 256, smem, stream>>>(__VA_ARGS__); break;           \
      case 1 * 8 + 2: KERNEL<1, 2><<<grid, 256, smem, stream>>>(__VA_ARGS__); break;           \
      case 1 * 8 + 3: case 1 * 8 + 4: KERNEL<1, 4><<<grid, 256, smem, stream>>>(__VA_ARGS__); break; \
      case 2 * 8 + 1: KERNEL<2, 1><<<grid, 256, smem, stream>>>(__VA_ARGS__); break;           \
      case 2 * 8 + 2: KERNEL<2, 2><<<grid, 256, smem, stream>>>(__VA_ARGS__); break;           \
      case 2 * 8 + 3: case 2 * 8 + 4: KERNEL<2, 4><<<grid, 256, smem, stream>>>(__VA_ARGS__); break; \
      case 3 * 8 + 1: KERNEL<3, 1><<<grid, 256, smem, stream>>>(__VA_ARGS__); break;           \
      case 3 * 8 + 2: KERNEL<3, 2><<<grid, 256, smem, stream>>>(__VA_ARGS__); break;           \
      case 3 * 8 + 3: case 3 * 8 + 4: KERNEL<3, 4><<<grid, 256, smem, stream>>>(__VA_ARGS__); break; \
      case 4 * 8 + 1: KERNEL<4, 1><<<grid, 256, smem, stream>>>(__VA_ARGS__); break;           \
      case 4 * 8 + 2: KERNEL<4, 2><<<grid, 256, smem, stream>>>(__VA_ARGS__); break;           \
      default: KERNEL<4, 4><<<grid, 256, smem, stream>>>(__VA_ARGS__); break;                  \
    }                                                                                       \
  } while (0)

}  // namespace

extern "C" {

int b200corr_altcorr_forward(const float *fmap1, const float *fmap2, const float *coords,
                             float *corr, int B, int N, int H1, int W1, int H2, int W2, int C,
                             int radius, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int e = check_alt("altcorr_forward", B, N, H1, W1, H2, W2, C, radius)) return e;
  if (B == 0) return 0;
  B200_CHECK(fmap1 && fmap2 && coords && corr, "altcorr_forward: null pointer");
  B200_CHECK((((uintptr_t)fmap1 | (uintptr_t)fmap2) & 15) == 0, "altcorr_forward: feature maps must be 16-byte aligned");
  dim3 grid((H1 * W1 + 7) / 8, N, B);
  ALT_DISPATCH(altcorr_fwd_kernel, 0, fmap1, fmap2, coords, corr, B, N, H1, W1, H2, W2, C);
  B200_LAUNCH_OK("altcorr_fwd_kernel");
  return 0;
}

int b200corr_altcorr_backward(const float *fmap1, const float *fmap2, const float *coords,
                              const float *corr_grad, float *fmap1_grad, float *fmap2_grad,
                              float *coords_grad, int B, int N, int H1, int W1, int H2, int W2,
                              int C, int radius, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int e = check_alt("altcorr_backward", B, N, H1, W1, H2, W2, C, radius)) return e;
  if (B == 0) return 0;
  B200_CHECK(fmap1 && fmap2 && coords && corr_grad && fmap1_grad && fmap2_grad,
             "altcorr_backward: null pointer");
  B200_CHECK((((uintptr_t)fmap1 | (uintptr_t)fmap2 | (uintptr_t)fmap1_grad | (uintptr_t)fmap2_grad) & 15) == 0,
             "altcorr_backward: feature maps and gradients must be 16-byte aligned");
  B200_CUDA(cudaMemsetAsync(fmap2_grad, 0, sizeof(float) * (size_t)B * H2 * W2 * C, stream));
  if (coords_grad)  // the reference allocates zeros and never writes them (correlation_kernel.cu:307)
    B200_CUDA(cudaMemsetAsync(coords_grad, 0, sizeof(float) * (size_t)B * N * H1 * W1 * 2, stream));
  dim3 grid((H1 * W1 + 7) / 8, 1, B);
  ALT_DISPATCH(altcorr_bwd_kernel, 0, fmap1, fmap2, coords, corr_grad, fmap1_grad, fmap2_grad, B, N, H1,
               W1, H2, W2, C);
  B200_LAUNCH_OK("altcorr_bwd_kernel");
  return 0;
}

}  // extern "C"
