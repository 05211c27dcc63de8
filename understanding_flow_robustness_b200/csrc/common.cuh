// common.cuh -- shared host/device helpers of libb200corr (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "b200corr.h"

// ---------------------------------------------------------------------------------- host side
namespace b200 {

void set_error(const char *fmt, ...);  // thread-local message for b200corr_last_error()
void count_launch(int n = 1);

#define B200_CHECK(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      b200::set_error(__VA_ARGS__);  \
      return -1;                     \
    }                                \
  } while (0)

#define B200_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t err__ = (expr);                                                           \
    if (err__ != cudaSuccess) {                                                           \
      b200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, \
                      __LINE__);                                                          \
      return -2;                                                                          \
    }                                                                                     \
  } while (0)

// launch check: catches configuration errors (bad smem size, bad grid) right at the call
#define B200_LAUNCH_OK(name)                                                               \
  do {                                                                                     \
    cudaError_t err__ = cudaGetLastError();                                                \
    if (err__ != cudaSuccess) {                                                            \
      b200::set_error("launch of %s failed: %s", name, cudaGetErrorString(err__));         \
      return -3;                                                                           \
    }                                                                                      \
    b200::count_launch();                                                                  \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

int num_sms();  // SM count of the current device (cached)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device and serialises with running work:
// call it once per (kernel, device).  `done` is a zero-initialised per-kernel flag array.
int set_max_smem_once(const void *kernel, int bytes, bool (&done)[64]);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda at link time).
// dims/strides innermost first; strides in BYTES for dims 1..rank-1 (dim 0 is contiguous).
int make_tensor_map(CUtensorMap *map, CUtensorMapDataType dtype, int rank, const void *base,
                    const uint64_t *dims, const uint64_t *strides_bytes, const uint32_t *box,
                    CUtensorMapSwizzle swizzle, CUtensorMapL2promotion l2promo,
                    const uint32_t *elem_strides = nullptr);

// ---- spatial correlation sampler back ends (q = the reference's 12 integers: kH kW patchH patchW
// padH padW dilationH dilationW dilation_patchH dilation_patchW dH dW)
int sampler_generic_forward(const void *in1, const void *in2, void *out, int B, int C, int H, int W,
                            int oH, int oW, const int *q, int dtype, cudaStream_t stream);
int sampler_generic_backward(const void *in1, const void *in2, const void *gout, void *gin1,
                             void *gin2, int B, int C, int H, int W, int oH, int oW, const int *q,
                             int dtype, cudaStream_t stream);
bool sampler_fast_applicable(int B, int C, int H, int W, const int *q, int dtype, int backward);
int sampler_fast_forward(const float *in1, const float *in2, float *out, int B, int C, int H, int W,
                         const int *q, cudaStream_t stream);
int sampler_fast_forward_merge(const float *in1, const float *in2, float *out, int B, int C, int H, int W,
                               const int *q, long long out_bstride, float slope, cudaStream_t stream);
int sampler_fast_backward(const float *in1, const float *in2, const float *gout, float *gin1,
                          float *gin2, int B, int C, int H, int W, const int *q, const int *plan,
                          cudaStream_t stream);
size_t sampler_fast_backward_plan_ints(int B, int C, int H, int W, const int *q);
int sampler_fast_backward_plan(int B, int C, int H, int W, const int *q, int *h_plan, size_t bytes);

}  // namespace b200

// -------------------------------------------------------------------------------- device side
#ifdef __CUDACC__
namespace b200dev {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe of a phase (try_wait may suspend the thread for a while; test_wait never does)
__device__ __forceinline__ bool mbar_test_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- TMA (cp.async.bulk.tensor), tile mode, global -> shared::cta, completion on an mbarrier
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *m, uint64_t *bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *m, uint64_t *bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void *dst, const CUtensorMap *m, uint64_t *bar, int c0,
                                            int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
      "r"(c4)
      : "memory");
}
// 1-D bulk copy global -> shared (no tensor map); bytes % 16 == 0, both addresses 16-B aligned
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, uint32_t bytes,
                                             uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ float4 lds128(const float *p) {
  return *reinterpret_cast<const float4 *>(p);
}

}  // namespace b200dev
#endif
