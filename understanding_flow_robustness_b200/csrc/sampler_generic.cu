// sampler_generic.cu -- spatial correlation sampler for ARBITRARY hyper-parameters and fp32 / fp64 /
// fp16 / bf16 storage.
//
// Replaces the role of correlation_cuda_kernel.cu:22-233 (reference CUDA kernels) for every
// configuration the register-blocked TMA kernels in sampler_fast.cu do not cover (kernel_size > 1,
// stride > 1, padding, dilation, fp64 for gradcheck-style tests).  Semantics follow the reference
// CPU loops correlation.cpp:9-73 exactly:
//   * forward : one thread per output element, accumulation order c -> i -> j;
//   * backward: gather form (no atomics), one thread per input element, contributions visited in
//               the order (ph, pw, h, w) in which the reference's scatter loops reach that element.
// Products and sums use __fmul_rn/__fadd_rn (no FMA contraction), so on identical inputs the result
// is bit-identical to the reference's CPU build -- the -m gpu tests assert equality, and the fast
// kernels are checked against these at full problem sizes.
// Half precision (the reference's CUDA dispatch accepts at::Half, correlation_cuda_kernel.cu:262,297;
// its CPU dispatch does not): fp16 / bf16 tensors are read and written in their own type, products
// and sums are carried in fp32 and rounded once on the final store -- at least as accurate as the
// reference, which accumulates in half.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

struct SamplerParams {
  int B, C, H, W, oH, oW;
  int kH, kW, patchH, patchW, padH, padW, dilH, dilW, dpH, dpW, sH, sW;
};

// storage type -> arithmetic type
template <typename T> struct Acc { using type = T; };
template <> struct Acc<__half> { using type = float; };
template <> struct Acc<__nv_bfloat16> { using type = float; };
template <typename T>
__device__ __forceinline__ typename Acc<T>::type ld(const T *p) { return *p; }
template <>
__device__ __forceinline__ float ld<__half>(const __half *p) { return __half2float(*p); }
template <>
__device__ __forceinline__ float ld<__nv_bfloat16>(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
template <typename T>
__device__ __forceinline__ void st(T *p, typename Acc<T>::type v) { *p = v; }
template <>
__device__ __forceinline__ void st<__half>(__half *p, float v) { *p = __float2half_rn(v); }
template <>
__device__ __forceinline__ void st<__nv_bfloat16>(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

template <typename T>
__device__ __forceinline__ T mul_rn(T a, T b);
template <>
__device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <>
__device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <typename T>
__device__ __forceinline__ T add_rn(T a, T b);
template <>
__device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
template <>
__device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }

// out[n, ph, pw, h, w]; consecutive threads -> consecutive w (coalesced on in1, in2 and out).
template <typename T>
__global__ void __launch_bounds__(256)
sampler_generic_forward_kernel(const T *__restrict__ in1, const T *__restrict__ in2,
                               T *__restrict__ out, SamplerParams p, long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int w = (int)(idx % p.oW);
    long long t = idx / p.oW;
    int h = (int)(t % p.oH);
    t /= p.oH;
    int pw = (int)(t % p.patchW);
    t /= p.patchW;
    int ph = (int)(t % p.patchH);
    int n = (int)(t / p.patchH);
    const int shiftU = (ph - (p.patchH - 1) / 2) * p.dpH;
    const int shiftV = (pw - (p.patchW - 1) / 2) * p.dpW;
    const int u = -p.padH + h * p.sH, v = -p.padW + w * p.sW;
    const T *a = in1 + (size_t)n * p.C * p.H * p.W;
    const T *b = in2 + (size_t)n * p.C * p.H * p.W;
    using A = typename Acc<T>::type;
    A acc = A(0);
    for (int c = 0; c < p.C; ++c) {
      for (int i = 0; i < p.kH; ++i) {
        int i1 = u + i * p.dilH, i2 = i1 + shiftU;
        if (i1 < 0 || i1 >= p.H || i2 < 0 || i2 >= p.H) continue;
        for (int j = 0; j < p.kW; ++j) {
          int j1 = v + j * p.dilW, j2 = j1 + shiftV;
          if (j1 < 0 || j1 >= p.W || j2 < 0 || j2 >= p.W) continue;
          acc = add_rn(acc, mul_rn(ld(a + (size_t)i1 * p.W + j1), ld(b + (size_t)i2 * p.W + j2)));
        }
      }
      a += (size_t)p.H * p.W;
      b += (size_t)p.H * p.W;
    }
    st(out + idx, acc);
  }
}

// WHICH == 1: grad_in1[n,c,y,x] = sum gout[n,ph,pw,h,w] * in2[n,c,y+dy,x+dx]   (y = tap of in1)
// WHICH == 2: grad_in2[n,c,y,x] = sum gout[n,ph,pw,h,w] * in1[n,c,y-dy,x-dx]   (y = tap of in2)
template <typename T, int WHICH>
__global__ void __launch_bounds__(256)
sampler_generic_backward_kernel(const T *__restrict__ other, const T *__restrict__ gout,
                                T *__restrict__ gin, SamplerParams p, long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int x = (int)(idx % p.W);
    long long t = idx / p.W;
    int y = (int)(t % p.H);
    t /= p.H;
    int c = (int)(t % p.C);
    int n = (int)(t / p.C);
    using A = typename Acc<T>::type;
    const T *oth = other + ((size_t)n * p.C + c) * p.H * p.W;
    A acc = A(0);
    for (int ph = 0; ph < p.patchH; ++ph) {
      const int dy = (ph - (p.patchH - 1) / 2) * p.dpH;
      const int y1 = (WHICH == 1) ? y : y - dy;  // tap position in input1
      const int y2 = y1 + dy;                    // tap position in input2
      if (y1 < 0 || y1 >= p.H || y2 < 0 || y2 >= p.H) continue;
      for (int pw = 0; pw < p.patchW; ++pw) {
        const int dx = (pw - (p.patchW - 1) / 2) * p.dpW;
        const int x1 = (WHICH == 1) ? x : x - dx;
        const int x2 = x1 + dx;
        if (x1 < 0 || x1 >= p.W || x2 < 0 || x2 >= p.W) continue;
        const A ov = (WHICH == 1) ? ld(oth + (size_t)y2 * p.W + x2) : ld(oth + (size_t)y1 * p.W + x1);
        const T *g = gout + (((size_t)n * p.patchH + ph) * p.patchW + pw) * p.oH * p.oW;
        // y1 = h*sH - padH + i*dilH ; i descending <=> h ascending (reference visiting order)
        for (int i = p.kH - 1; i >= 0; --i) {
          int hn = y1 + p.padH - i * p.dilH;
          if (hn < 0 || hn % p.sH) continue;
          int h = hn / p.sH;
          if (h >= p.oH) continue;
          for (int j = p.kW - 1; j >= 0; --j) {
            int wn = x1 + p.padW - j * p.dilW;
            if (wn < 0 || wn % p.sW) continue;
            int w = wn / p.sW;
            if (w >= p.oW) continue;
            acc = add_rn(acc, mul_rn(ld(g + (size_t)h * p.oW + w), ov));
          }
        }
      }
    }
    st(gin + idx, acc);
  }
}

template <typename T>
int launch_forward(const void *in1, const void *in2, void *out, const SamplerParams &p,
                   cudaStream_t stream) {
  long long total = (long long)p.B * p.patchH * p.patchW * p.oH * p.oW;
  if (total == 0) return 0;
  long long blocks = (total + 255) / 256;
  long long cap = (long long)b200::num_sms() * 32;
  int grid = (int)(blocks < cap ? blocks : cap);
  sampler_generic_forward_kernel<T><<<grid, 256, 0, stream>>>((const T *)in1, (const T *)in2,
                                                              (T *)out, p, total);
  B200_LAUNCH_OK("sampler_generic_forward_kernel");
  return 0;
}

template <typename T>
int launch_backward(const void *in1, const void *in2, const void *gout, void *gin1, void *gin2,
                    const SamplerParams &p, cudaStream_t stream) {
  long long total = (long long)p.B * p.C * p.H * p.W;
  if (total == 0) return 0;
  long long blocks = (total + 255) / 256;
  long long cap = (long long)b200::num_sms() * 32;
  int grid = (int)(blocks < cap ? blocks : cap);
  sampler_generic_backward_kernel<T, 1><<<grid, 256, 0, stream>>>((const T *)in2, (const T *)gout,
                                                                  (T *)gin1, p, total);
  B200_LAUNCH_OK("sampler_generic_backward_kernel<1>");
  sampler_generic_backward_kernel<T, 2><<<grid, 256, 0, stream>>>((const T *)in1, (const T *)gout,
                                                                  (T *)gin2, p, total);
  B200_LAUNCH_OK("sampler_generic_backward_kernel<2>");
  return 0;
}

}  // namespace

namespace b200 {

int sampler_generic_forward(const void *in1, const void *in2, void *out, int B, int C, int H, int W,
                            int oH, int oW, const int *q, int dtype, cudaStream_t stream) {
  SamplerParams p{B, C, H, W, oH, oW, q[0], q[1], q[2], q[3], q[4], q[5],
                  q[6], q[7], q[8], q[9], q[10], q[11]};
  switch (dtype) {
    case B200CORR_F64: return launch_forward<double>(in1, in2, out, p, stream);
    case B200CORR_F16: return launch_forward<__half>(in1, in2, out, p, stream);
    case B200CORR_BF16: return launch_forward<__nv_bfloat16>(in1, in2, out, p, stream);
    default: return launch_forward<float>(in1, in2, out, p, stream);
  }
}

int sampler_generic_backward(const void *in1, const void *in2, const void *gout, void *gin1,
                             void *gin2, int B, int C, int H, int W, int oH, int oW, const int *q,
                             int dtype, cudaStream_t stream) {
  SamplerParams p{B, C, H, W, oH, oW, q[0], q[1], q[2], q[3], q[4], q[5],
                  q[6], q[7], q[8], q[9], q[10], q[11]};
  switch (dtype) {
    case B200CORR_F64: return launch_backward<double>(in1, in2, gout, gin1, gin2, p, stream);
    case B200CORR_F16: return launch_backward<__half>(in1, in2, gout, gin1, gin2, p, stream);
    case B200CORR_BF16: return launch_backward<__nv_bfloat16>(in1, in2, gout, gin1, gin2, p, stream);
    default: return launch_backward<float>(in1, in2, gout, gin1, gin2, p, stream);
  }
}

}  // namespace b200
