// common.cu -- error reporting, launch counting, tensor-map encoding, FP32-pipe peak probe.
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace b200 {

static thread_local char g_error[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

int set_max_smem_once(const void *kernel, int bytes, bool (&done)[64]) {
  int dev = 0;
  B200_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !done[dev]) {
    B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (dev >= 0 && dev < 64) done[dev] = true;
  }
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tensor_map(CUtensorMap *map, CUtensorMapDataType dtype, int rank, const void *base,
                    const uint64_t *dims, const uint64_t *strides_bytes, const uint32_t *box,
                    CUtensorMapSwizzle swizzle, CUtensorMapL2promotion l2promo,
                    const uint32_t *elem_strides) {
  EncodeTiledFn fn = get_encode_fn();
  B200_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the CUDA driver");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = elem_strides ? elem_strides[i] : 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i];
  }
  CUresult r = fn(map, dtype, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, l2promo,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);  // out-of-bounds box elements read as 0
  if (r != CUDA_SUCCESS) {
    set_error(
        "cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu %llu %llu %llu %llu] box "
        "[%u %u %u %u %u] base %p",
        (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
        (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
        (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], rank > 1 ? box[1] : 0,
        rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0, base);
    return -4;
  }
  return 0;
}

}  // namespace b200

// ------------------------------------------------------------------------------ FP32 peak probe
// 8 independent FMA chains per thread, operands in registers: the shape the sampler kernels issue.
__global__ void __launch_bounds__(256) fp32_peak_kernel(float *sink, int iters, float seed) {
  float a0 = seed + threadIdx.x, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
  float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  float x = 1.0000001f + seed, y = 0.9999999f - seed;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      a0 = fmaf(a0, x, y);
      a1 = fmaf(a1, y, x);
      a2 = fmaf(a2, x, y);
      a3 = fmaf(a3, y, x);
      a4 = fmaf(a4, x, y);
      a5 = fmaf(a5, y, x);
      a6 = fmaf(a6, x, y);
      a7 = fmaf(a7, y, x);
    }
  }
  float s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 12345.678f) sink[0] = s;  // never true in practice; keeps the chains alive
}

extern "C" {

int b200corr_version(void) { return B200CORR_VERSION; }

const char *b200corr_last_error(void) { return b200::g_error; }

uint64_t b200corr_launch_count(void) { return b200::g_launches.load(std::memory_order_relaxed); }

int b200corr_measure_fp32_peak(int iters, float *tflops, void *stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  B200_CHECK(iters > 0 && tflops != nullptr, "measure_fp32_peak: bad arguments");
  float *sink = nullptr;
  B200_CUDA(cudaMalloc(&sink, sizeof(float)));  // diagnostics only: the data path never allocates
  cudaEvent_t e0, e1;
  B200_CUDA(cudaEventCreate(&e0));
  B200_CUDA(cudaEventCreate(&e1));
  const int blocks = b200::num_sms() * 8, threads = 256;
  fp32_peak_kernel<<<blocks, threads, 0, stream>>>(sink, 16, 0.f);  // warm-up
  float best_ms = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    B200_CUDA(cudaEventRecord(e0, stream));
    fp32_peak_kernel<<<blocks, threads, 0, stream>>>(sink, iters, 0.f);
    B200_CUDA(cudaEventRecord(e1, stream));
    B200_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    B200_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best_ms) best_ms = ms;
  }
  B200_LAUNCH_OK("fp32_peak_kernel");
  b200::count_launch(5);
  const double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
  *tflops = (float)(flops / (best_ms * 1e-3) / 1e12);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  return 0;
}

}  // extern "C"
